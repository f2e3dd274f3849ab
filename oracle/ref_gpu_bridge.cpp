/* oracle/ref_gpu_bridge.cpp -- TEST INFRASTRUCTURE ONLY (the drop-in demonstration).
 *
 * The UNMODIFIED reference host program -- command line, Pindel-text/BAM-config input, window loop, SV drivers,
 * reporter, _D/_SI/_INV/_TD writers -- linked against libpindel_pg.so.  This file supplies the two seam functions with
 * the reference's own signatures,
 *
 *     void GetCloseEnd(const std::string&, SPLIT_READ&)                          src/pindel.h:523
 *     void SearchFarEnds(const std::string&, std::vector<SPLIT_READ>&,
 *                        const Chromosome&)                                      src/pindel.h:741
 *
 * in the executable; the reference objects live in libpindel_refcore.so (built -fPIC from the untouched sources), so the
 * dynamic linker binds every call the reference makes to these two names (reader.cpp:254,304, read_buffer.cpp:47,
 * pindel.cpp:1888) to the definitions below.  No reference source is patched or copied.
 *
 * Besides the two seam functions the bridge supplies the two places where the reference *batches* reads before calling
 * GetCloseEnd -- ReadBuffer::flush for BAM input (read_buffer.cpp:36-101) and ReadInRead for Pindel-text input
 * (reader.cpp:196-361), both reachable by symbol -- as the batched binding INTEGRATION.md describes: one pg_close_end
 * call per batch.  The per-read GetCloseEnd below remains for any other caller; SearchFarEnds is a batch call anyway.
 * Set PG_BRIDGE_UNBATCHED=1 to make the text path go through the per-read shim instead (the reference's own ReadInRead).
 *
 * Row N1 of SURVEY.md section 8(f): the seven SV drivers SearchSVs calls (src/pindel.cpp:1140-1164) are supplied here too --
 * SearchVariant::Search (D and SI), searchIndels, searchTandemDuplications(NT), searchInversions(NT).  The first driver of
 * a window sends every read with its UniquePoint lists to pg_classify once (all drivers in one launch); each driver then
 * replays its own records (SPLIT_READ fields, Vars[box], FutureReads_SR) and calls the reference's reporter
 * (outputResults / SortOutput*) exactly as before.  PG_BRIDGE_CPU_SV=1 is not offered: there is no CPU path behind the ABI.
 */
#include "pindel.h"
#include "control_state.h"
/* BDData::loadRegion is supplied below (row N2): a member function, so this TU -- test infrastructure -- sees the members it
 * works on (the event list UpdateBD filled, the window, the mask) as public.  A maintainer would simply edit the member. */
#define private public
#include "bddata.h"
#undef private
#include "read_buffer.h"
#include "reader.h"
#include "pindel_read_reader.h"
#include "logstream.h"
#include "user_defined_settings.h"
#include "control_state.h"
#include "logdef.h"
#include "reporter.h"
#include "search_variant.h"
#include "searchdeletions.h"
#include "search_deletions_nt.h"
#include "search_tandem_duplications.h"
#include "search_tandem_duplications_nt.h"
#include "search_inversions.h"
#include "search_inversions_nt.h"

#include <cstdio>
#include <dlfcn.h>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../include/pindel_pg.h"

extern int pindel_reference_main(int argc, char *argv[]);
extern BDData g_bdData;
extern UserDefinedSettings *userSettings;
extern short g_MinClose;

namespace {
pg_ctx *g_ctx = NULL;
pg_group *g_group = NULL;      /* PG_BRIDGE_DEVICES > 1 */
std::mutex g_mutex;
/* where the drop-in's wall time goes (printed at exit): per binding the host-side marshalling, the library call (copies +
 * kernels) and the host-side application of the results */
struct Acc { double s; unsigned long n; };
Acc t_close_marshal = {0, 0}, t_close_lib = {0, 0}, t_close_apply = {0, 0}, t_far_marshal = {0, 0}, t_far_lib = {0, 0}, t_far_apply = {0, 0},
    t_sv_lib = {0, 0}, t_sv_total = {0, 0}, t_box_lib = {0, 0}, t_rp_lib = {0, 0}, t_inflate_lib = {0, 0}, t_index_lib = {0, 0}, t_rpd_lib = {0, 0},
    t_sr_lib = {0, 0}, t_sr_host = {0, 0}, t_setup = {0, 0}, t_bd_lib = {0, 0};
struct Tick {
   Acc &a;
   std::chrono::steady_clock::time_point t0;
   explicit Tick(Acc &acc) : a(acc), t0(std::chrono::steady_clock::now()) {}
   ~Tick() { a.s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); a.n++; }
};
const std::chrono::steady_clock::time_point g_t_start = std::chrono::steady_clock::now();
unsigned long g_close_calls = 0, g_far_reads = 0, g_sv_reads = 0, g_sv_records = 0, g_sv_resident = 0, g_rp_rows = 0;
bool g_bdOnDevice = false;                     /* the current window's BreakDancer region is loaded in the library (pg_bd_load_region) */
unsigned long g_bd_regions = 0, g_bd_events = 0;
bool g_flushPresearched = false;               /* the reads in the ReadBuffer already went through GetCloseEnd on the device */
const SPLIT_READ *g_farResidentBase = NULL;   /* Reads_SR of the last SearchFarEnds call, still resident on the device */
size_t g_farResidentN = 0;

void die(const char *what)
{
   std::fprintf(stderr, "pindel_gpu_demo: %s: %s\n", what, g_ctx ? pg_last_error(g_ctx) : pg_create_error());
   std::exit(EXIT_FAILURE);   /* the reference's own error convention: message + exit (e.g. src/pindel.cpp:674) */
}

void ensureContext()
{
   if (g_ctx) return;
   Tick tk(t_setup);
   pg_settings s;
   pg_default_settings(&s);
   s.seq_error_rate = userSettings->Seq_Error_Rate;
   s.sensitivity = userSettings->sensitivity;
   s.max_mismatch_rate = userSettings->MaximumAllowedMismatchRate;
   s.additional_mismatch = userSettings->ADDITIONAL_MISMATCH;
   s.min_close = g_MinClose;
   s.min_perfect_match = userSettings->Min_Perfect_Match_Around_BP;
   s.max_range_index = userSettings->MaxRangeIndex;
   /* PG_BRIDGE_DEVICES=N: the two batched seams fan out over N GPUs inside the library (pg_group); everything else -- and the
    * classification, which then takes the host arrays -- runs on the group's first device */
   std::vector<int32_t> devs;        /* "N" = devices 0 .. N-1, or a comma list of ordinals ("0,0,0": three contexts on one GPU) */
   if (const char *e = getenv("PG_BRIDGE_DEVICES")) {
      if (strchr(e, ',')) { for (const char *p = e; *p; ) { devs.push_back(atoi(p)); p = strchr(p, ','); if (!p) break; p++; } }
      else for (int k = 0; k < atoi(e); k++) devs.push_back(k);
   }
   const int n_dev = (int)devs.size();
   if (n_dev > 1) {
      g_group = pg_group_create(&s, n_dev, &devs[0]);
      if (!g_group) die("pg_group_create");
      g_ctx = pg_group_context(g_group, 0);
   } else {
      g_ctx = pg_create(&s);
   }
   if (!g_ctx) die("pg_create");
   std::vector<const char *> seqs;
   std::vector<uint32_t> lens;
   for (unsigned i = 0; g_genome.getChr(i) != NULL; i++) {
      seqs.push_back(g_genome.getChr(i)->getSeq().c_str());
      lens.push_back(g_genome.getChr(i)->getCompSize());
   }
   if (g_group) {
      if (pg_group_set_reference(g_group, (int)seqs.size(), &seqs[0], &lens[0]) != PG_OK) { std::fprintf(stderr, "pindel_gpu_demo: %s\n", pg_group_last_error(g_group)); std::exit(EXIT_FAILURE); }
   } else if (pg_set_reference(g_ctx, (int)seqs.size(), &seqs[0], &lens[0]) != PG_OK) die("pg_set_reference");
   /* the library derives g_maxMismatch itself; it must agree with the host's table */
   std::vector<uint32_t> tab(500);
   pg_max_mismatch_table(g_ctx, &tab[0]);
   for (unsigned i = 0; i < 500 && i < g_maxMismatch.size(); i++)
      if (tab[i] != g_maxMismatch[i]) { std::fprintf(stderr, "pindel_gpu_demo: mismatch table differs at %u\n", i); std::exit(EXIT_FAILURE); }
}

void appendRuns(SortedUniquePoints &dst, const pg_run *runs, unsigned n)
{
   for (unsigned r = 0; r < n; r++) {
      const pg_run &q = runs[r];
      const Chromosome *c = g_genome.getChr((unsigned)q.chr);
      for (unsigned j = 0; j < q.count; j++) {
         unsigned absLoc = (q.dirstrand & 1u) ? q.abs_loc0 - j : q.abs_loc0 + j;
         dst.push_back(UniquePoint(c, (short)(q.len0 + j), absLoc, (q.dirstrand & 1u) ? BACKWARD : FORWARD,
                                   (q.dirstrand & 2u) ? ANTISENSE : SENSE, (short)q.mm));
      }
   }
}
} // namespace

/* pg_close_end over reads[idx[0..m)], results applied the way GetCloseEnd leaves a SPLIT_READ. */
static void closeEndBatch(std::vector<SPLIT_READ> &reads, const std::vector<int> &idx)
{
   const int m = (int)idx.size();
   if (m == 0) return;
   Tick *tk = new Tick(t_close_marshal);
   std::string blob;
   std::vector<uint32_t> off(m + 1, 0), pos(m);
   std::vector<char> md(m);
   std::vector<int16_t> isz(m);
   std::vector<int32_t> chr(m);
   for (int k = 0; k < m; k++) {
      const SPLIT_READ &r = reads[idx[k]];
      blob += r.getUnmatchedSeq();
      off[k + 1] = (uint32_t)blob.size();
      md[k] = r.MatchedD; pos[k] = r.MatchedRelPos; isz[k] = r.InsertSize; chr[k] = g_genome.getChrID(r.FragName);
      if (r.InsertSize > g_maxInsertSize) g_maxInsertSize = r.InsertSize;     /* side effect of GetCloseEndInner, src/pindel.cpp:2257 */
   }
   pg_reads in;
   memset(&in, 0, sizeof(in));
   in.n = m; in.seq_blob = blob.c_str(); in.seq_off = &off[0]; in.matched_d = &md[0]; in.matched_rel_pos = &pos[0];
   in.insert_size = &isz[0]; in.chr_id = &chr[0];
   std::vector<pg_read_out> out(m);
   std::vector<pg_run> runs((size_t)m * 16 + 1024);
   pg_results res;
   res.out = &out[0]; res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size(); res.n_runs = 0;
   delete tk; tk = new Tick(t_close_lib);
   int rc = g_group ? pg_group_close_end(g_group, &in, &res) : pg_close_end(g_ctx, &in, &res);
   if (rc == PG_ERR_RUNS_CAP) {
      runs.resize(res.n_runs);
      res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size();
      rc = g_group ? pg_group_close_end(g_group, &in, &res) : pg_close_end(g_ctx, &in, &res);
   }
   if (rc != PG_OK) { if (g_group) std::fprintf(stderr, "pindel_gpu_demo: %s\n", pg_group_last_error(g_group)); die("pg_close_end (batch)"); }
   delete tk;
   Tick tk2(t_close_apply);
   for (int k = 0; k < m; k++) {
      SPLIT_READ &read = reads[idx[k]];
      if (out[k].flags & PG_READ_UNSUPPORTED) die("read not supported by the library");
      for (unsigned f = 0; f < PG_READ_NFLIPS(out[k].flags); f++)
         read.setUnmatchedSeq(ReverseComplement(read.getUnmatchedSeq()));
      read.UP_Close.clear();
      appendRuns(read.UP_Close, &runs[out[k].run_off], out[k].n_close_runs);
      if (!read.UP_Close.empty()) read.Used = false;
   }
   g_close_calls += (unsigned long)m;
}

/* ReadInRead, batched (replaces reader.cpp:196-361 by symbol): scan the Pindel-text file, keep the reads anchored in
 * the window, search their close ends in ONE batch instead of 1000-read OpenMP batches, then post-process them in file
 * order exactly as the reference does (reader.cpp:258-293 / 307-338). */
extern unsigned int g_NumReadInWindow, g_NumReadScanned, g_InWinPlus, g_InWinMinus, g_CloseMappedPlus, g_CloseMappedMinus;
short ReadInRead(PindelReadReader &inf_ReadSeq, const std::string &FragName, const std::string &CurrentChr,
                 std::vector<SPLIT_READ> &Reads, const SearchWindow &currentWindow)
{
   *logStream << "Scanning and processing reads anchored in " << FragName << std::endl;
   std::vector<SPLIT_READ> BufferReads;
   SPLIT_READ Temp_One_Read;
   inf_ReadSeq.Reset();
   while (inf_ReadSeq.HasNext()) {
      Temp_One_Read = inf_ReadSeq.NextRead();
      if (Temp_One_Read.Name[0] != FirstCharReadName) {
         *logStream << "Something wrong with the read name: " << Temp_One_Read.Name << std::endl;
         Reads.clear();
         return EXIT_FAILURE;
      }
      g_NumReadScanned++;
      if (Temp_One_Read.MatchedD != Minus && Temp_One_Read.MatchedD != Plus) {
         *logStream << Temp_One_Read.Name << std::endl << Temp_One_Read.getUnmatchedSeq() << std::endl << Temp_One_Read.MatchedD << " ..." << std::endl << "+/-" << std::endl;
         return EXIT_FAILURE;
      }
      if (Temp_One_Read.MatchedRelPos > g_maxPos) g_maxPos = Temp_One_Read.MatchedRelPos;
      if (Temp_One_Read.FragName == FragName && Temp_One_Read.MatchedRelPos >= currentWindow.getStart() &&
          Temp_One_Read.MatchedRelPos < currentWindow.getEnd()) {
         g_NumReadInWindow++;
         if (Temp_One_Read.MatchedD == Plus) g_InWinPlus++; else g_InWinMinus++;
         if (Temp_One_Read.MatchedRelPos > currentWindow.getChromosome()->getBiolSize())
            Temp_One_Read.MatchedRelPos = currentWindow.getChromosome()->getBiolSize();
         BufferReads.push_back(Temp_One_Read);
      }
   }
   if (std::getenv("PG_BRIDGE_UNBATCHED")) {
      for (size_t i = 0; i < BufferReads.size(); i++) GetCloseEnd(CurrentChr, BufferReads[i]);   /* one-read batches */
   } else {
      std::lock_guard<std::mutex> lock(g_mutex);
      ensureContext();
      std::vector<int> idx(BufferReads.size());
      for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
      closeEndBatch(BufferReads, idx);
   }
   int UPCLOSE_COUNTER = 0;
   for (size_t i = 0; i < BufferReads.size(); i++) {
      SPLIT_READ &currentRead = BufferReads[i];
      if (currentRead.UP_Close.size()) {
         UPCLOSE_COUNTER++;
         if (g_reportLength < currentRead.getReadLength()) g_reportLength = currentRead.getReadLength();
         currentRead.Used = false;
         currentRead.UniqueRead = true;
         CleanUniquePoints(currentRead.UP_Close);
         currentRead.CloseEndLength = currentRead.UP_Close[currentRead.UP_Close.size() - 1].LengthStr;
         if (currentRead.MatchedD == Plus)
            currentRead.LeftMostPos = currentRead.UP_Close[currentRead.UP_Close.size() - 1].AbsLoc + 1 - currentRead.CloseEndLength;
         else
            currentRead.LeftMostPos = currentRead.UP_Close[currentRead.UP_Close.size() - 1].AbsLoc + currentRead.CloseEndLength - currentRead.getReadLength();
         currentRead.SampleName2Number.insert(std::pair<std::string, unsigned>(currentRead.Tag, 1));
         Reads.push_back(currentRead);
         if (currentRead.MatchedD == Plus) g_CloseMappedPlus++; else g_CloseMappedMinus++;
         g_sampleNames.insert(currentRead.Tag);
      }
   }
   *logStream << "last one: " << BufferReads.size() << " and UPCLOSE= " << UPCLOSE_COUNTER << std::endl;
   if (FirstChr) FirstChr = false;
   *logStream << "Number of problematic reads in current window:            \t" << g_NumReadInWindow << ", + " << g_InWinPlus << " - " << g_InWinMinus << std::endl;
   *logStream << "Number of split-reads where the close end could be mapped:\t" << Reads.size() << ", + " << g_CloseMappedPlus << " - " << g_CloseMappedMinus << std::endl;
   if (Reads.size() == 0) return 0;
   *logStream << " finished!" << std::endl;
   return 0;
}

void GetCloseEnd(const std::string &CurrentChrSeq, SPLIT_READ &read)
{
   (void)CurrentChrSeq;
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   g_close_calls++;
   if (read.InsertSize > g_maxInsertSize) g_maxInsertSize = read.InsertSize;   /* side effect of GetCloseEndInner, src/pindel.cpp:2257 */
   const std::string &seq = read.getUnmatchedSeq();
   uint32_t off[2] = {0, (uint32_t)seq.size()};
   char md = read.MatchedD;
   uint32_t pos = read.MatchedRelPos;
   int16_t isz = read.InsertSize;
   int32_t chr = g_genome.getChrID(read.FragName);
   pg_reads in;
   memset(&in, 0, sizeof(in));
   in.n = 1; in.seq_blob = seq.c_str(); in.seq_off = off; in.matched_d = &md; in.matched_rel_pos = &pos;
   in.insert_size = &isz; in.chr_id = &chr;
   pg_read_out out;
   std::vector<pg_run> runs(1024);
   pg_results res;
   res.out = &out; res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size(); res.n_runs = 0;
   if (pg_close_end(g_ctx, &in, &res) != PG_OK) die("pg_close_end");
   if (out.flags & PG_READ_UNSUPPORTED) die("read not supported by the library");
   for (unsigned f = 0; f < PG_READ_NFLIPS(out.flags); f++)       /* replay the flips on the host copy, src/pindel.cpp:2545 */
      read.setUnmatchedSeq(ReverseComplement(read.getUnmatchedSeq()));
   read.UP_Close.clear();
   appendRuns(read.UP_Close, &runs[out.run_off], out.n_close_runs);
   if (!read.UP_Close.empty()) read.Used = false;                  /* src/pindel.cpp:2294, :2321 */
}

/* ReadBuffer::flush, batched (replaces read_buffer.cpp:36-101 by symbol; same observable behaviour: reads with a close
 * end are post-processed by updateReadAfterCloseEndMapping and appended to m_filteredReads in buffer order -- the
 * order the reference produces with one thread). */
void ReadBuffer::flush()
{
   if (!g_flushPresearched) {          /* reads from the device's split-read filter arrive with their close ends (pg_bam_sr_close_end) */
      std::lock_guard<std::mutex> lock(g_mutex);
      ensureContext();
      std::vector<int> idx;                  /* reads that take part (MapperSplit == false, read_buffer.cpp:46) */
      for (int i = 0; i < m_currentsize; i++) if (!m_rawreads[i].MapperSplit) idx.push_back(i);
      closeEndBatch(m_rawreads, idx);
   }
   for (int i = 0; i < m_currentsize; i++) {
      if (m_rawreads[i].hasCloseEnd()) {
         updateReadAfterCloseEndMapping(m_rawreads[i]);
         m_rawreads[i].SampleName2Number.insert(std::pair<std::string, unsigned>(m_rawreads[i].Tag, 1));
         m_filteredReads.push_back(m_rawreads[i]);
      }
   }
   m_rawreads.clear();
   m_currentsize = 0;
   std::cout << "The number of one end mapped read: " << m_filteredReads.size() << std::endl;
}

void SearchFarEnds(const std::string &chromosomeSeq, std::vector<SPLIT_READ> &reads, const Chromosome &currentChromosome)
{
   (void)chromosomeSeq;
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   const int n = (int)reads.size();
   Tick *tk = new Tick(t_far_marshal);
   std::string blob;
   std::vector<uint32_t> off(n + 1, 0), region_off(n + 1, 0);
   std::vector<int32_t> chr(n);
   std::vector<pg_close_info> ci(n);
   std::vector<pg_region> regions;
   std::vector<char> md(n);
   std::vector<uint32_t> pos(n);
   std::vector<int16_t> isz(n);
   bool staleFarEnd = false;   /* a read that is skipped here but carries an older UP_Far: the device copy would lack it */
   for (int i = 0; i < n; i++) {
      const SPLIT_READ &r = reads[i];
      blob += r.getUnmatchedSeq();
      off[i + 1] = (uint32_t)blob.size();
      chr[i] = g_genome.getChrID(r.FragName);
      md[i] = r.MatchedD; pos[i] = r.MatchedRelPos; isz[i] = r.InsertSize;   /* not used by the far-end search; the drivers need them */
      if ((r.MapperSplit || r.UP_Close.empty()) && !r.UP_Far.empty()) staleFarEnd = true;
      ci[i].pad = 0;
      if (r.MapperSplit || r.UP_Close.empty()) { ci[i].max_len = 0; ci[i].last_abs_loc = 0; }   /* src/pindel.cpp:1129 */
      else {
         ci[i].max_len = (uint16_t)r.UP_Close.MaxLen();
         ci[i].last_abs_loc = r.getLastAbsLocCloseEnd();
         if (g_bdOnDevice) continue;       /* the library looks every read's cluster up in the region loadRegion gave it */
         const SearchWindowCluster &cl = g_bdData.getCorrespondingSearchWindowCluster(r);      /* src/pindel.cpp:1006 */
         for (unsigned k = 0; k < cl.size(); k++) {
            pg_region g;
            g.chr = -1;
            for (unsigned c = 0; g_genome.getChr(c) != NULL; c++) if (g_genome.getChr(c) == cl[k].getChromosome()) g.chr = (int32_t)c;
            g.start = cl[k].getStart(); g.end = cl[k].getEnd();
            regions.push_back(g);
         }
      }
      if (!g_bdOnDevice) region_off[i + 1] = (uint32_t)regions.size();
   }
   (void)currentChromosome;
   pg_reads in;
   memset(&in, 0, sizeof(in));
   in.n = n; in.seq_blob = blob.c_str(); in.seq_off = &off[0]; in.chr_id = n ? &chr[0] : NULL; in.close_info = n ? &ci[0] : NULL;
   in.matched_d = n ? &md[0] : NULL; in.matched_rel_pos = n ? &pos[0] : NULL; in.insert_size = n ? &isz[0] : NULL;
   if (!regions.empty()) { in.regions = &regions[0]; in.region_off = &region_off[0]; }
   std::vector<pg_read_out> out(n ? n : 1);
   std::vector<pg_run> runs((size_t)n * 16 + 1024);
   pg_results res;
   res.out = &out[0]; res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size(); res.n_runs = 0;
   delete tk; tk = new Tick(t_far_lib);
   int rc = g_group ? pg_group_far_end(g_group, &in, &res) : pg_far_end(g_ctx, &in, &res);
   if (rc == PG_ERR_RUNS_CAP) {
      runs.resize(res.n_runs);
      res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size();
      rc = g_group ? pg_group_far_end(g_group, &in, &res) : pg_download(g_ctx, &res);
   }
   if (rc != PG_OK) { if (g_group) std::fprintf(stderr, "pindel_gpu_demo: %s\n", pg_group_last_error(g_group)); die("pg_far_end"); }
   delete tk;
   Tick tk2(t_far_apply);
   for (int i = 0; i < n; i++) {
      if (ci[i].max_len == 0) continue;
      reads[i].UP_Far.clear();
      appendRuns(reads[i].UP_Far, &runs[out[i].run_off + out[i].n_close_runs], out[i].n_far_runs);
      g_far_reads++;
   }
   /* the batch stays on the device: the first SV driver of this window classifies it there (pg_classify_resident) */
   g_farResidentBase = (n && !staleFarEnd && !g_group) ? &reads[0] : NULL;      /* a group's batch is spread over its devices */
   g_farResidentN = (size_t)n;
   *logStream << "Far end searching completed for this window." << std::endl;   /* src/pindel.cpp:1137 */
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Row N2, last part: BDData::loadRegion (src/bddata.cpp:852-946), called once per window (src/pindel.cpp:1853).  The reference
 * sweeps every position of the padded window to fill a position -> cluster mask that getCorrespondingSearchWindowCluster
 * (:949-967) then reads once per split read inside SearchFarEnd (src/pindel.cpp:1006).  Here the window's events go to the
 * library (pg_bd_load_region) and pg_far_end derives every read's cluster on the device; SearchFarEnds above then sends no
 * clusters.  The host keeps what the other readers of the state expect: the padded window and the empty cluster 0.  The
 * reference's own function runs instead (dlsym(RTLD_NEXT)) with PG_BRIDGE_HOST_BD=1, when -Q/-b ask for isBreakDancerEvent
 * (which reads the mask itself, src/bddata.cpp:981-995), or when an event names a chromosome that is not loaded (the
 * reference's error path, src/control_state.cpp:62-70).
 * ------------------------------------------------------------------------------------------------------------------- */
extern bool doOutputBreakdancerEvents();       /* src/pindel.cpp:521 */
void BDData::loadRegion(const SearchWindow &searchWindow)
{
   typedef void (*Fn)(BDData *, const SearchWindow &);
   static Fn original = (Fn)dlsym(RTLD_NEXT, "_ZN6BDData10loadRegionERK12SearchWindow");
   static const bool hostBD = getenv("PG_BRIDGE_HOST_BD") != NULL;
   std::lock_guard<std::mutex> lock(g_mutex);
   g_bdOnDevice = false;
   bool device = !hostBD && !doOutputBreakdancerEvents() && this == &g_bdData;
   std::vector<pg_bd_event> ev;
   std::vector<int32_t> rank;
   int32_t chrId = -1;
   if (device) {
      /* chromosome index by name; rank of every loaded chromosome's name in string order (the comparators compare names) */
      std::map<std::string, int32_t> idOf;
      std::vector<std::string> names;
      for (unsigned c = 0; g_genome.getChr(c) != NULL; c++) { idOf[g_genome.getChr(c)->getName()] = (int32_t)c; names.push_back(g_genome.getChr(c)->getName()); }
      rank.resize(names.size());
      { int32_t k = 0; for (std::map<std::string, int32_t>::const_iterator it = idOf.begin(); it != idOf.end(); ++it) rank[it->second] = k++; }
      const std::string &chrName = searchWindow.getChromosome()->getName();
      chrId = idOf.count(chrName) ? idOf[chrName] : -1;
      if (chrId < 0 || idOf.size() != names.size()) device = false;
      /* only the window chromosome's events can fall between the two bounds of the sorted list (:879-885); they keep their order */
      for (size_t i = 0; device && i < m_bdEvents.size(); i++) {
         const BreakDancerEvent &e = m_bdEvents[i];
         if (e.first.getChromosomeName() != chrName) continue;
         std::map<std::string, int32_t>::const_iterator b = idOf.find(e.second.getChromosomeName());
         if (b == idOf.end()) { device = false; break; }
         pg_bd_event x;
         x.chr_a = chrId; x.pos_a = e.first.position; x.pos2_a = e.first.position2;
         x.chr_b = b->second; x.pos_b = e.second.position; x.pos2_b = e.second.position2;
         ev.push_back(x);
      }
   }
   if (device) {
      ensureContext();
      Tick tk(t_bd_lib);
      const int nctx = g_group ? pg_group_size(g_group) : 1;
      for (int k = 0; k < nctx && device; k++) {
         pg_ctx *c = g_group ? pg_group_context(g_group, k) : g_ctx;
         if (pg_bd_load_region(c, ev.empty() ? NULL : &ev[0], (uint32_t)ev.size(), &rank[0], chrId, searchWindow.getStart(), searchWindow.getEnd()) != PG_OK) {
            std::fprintf(stderr, "pindel_gpu_demo: pg_bd_load_region: %s -- the reference's loadRegion runs\n", pg_last_error(c));
            device = false;
         }
      }
   }
   if (!device) {
      if (g_ctx) {
         const int nctx = g_group ? pg_group_size(g_group) : 1;
         for (int k = 0; k < nctx; k++) pg_bd_unload(g_group ? pg_group_context(g_group, k) : g_ctx);
      }
      original(this, searchWindow);
      return;
   }
   /* host state as the reference leaves it, minus the mask (:856-872, :895-899) */
   m_currentWindow = searchWindow;
   m_currentWindow.setStart(m_currentWindow.getStart() >= 3000 ? m_currentWindow.getStart() - 3000 : 0);
   m_currentWindow.setEnd(m_currentWindow.getEnd() + 3000);
   delete[] m_breakDancerMask;
   m_breakDancerMask = NULL;
   m_regionsToScanCollection.clear();
   m_regionsToScanCollection.push_back(SearchWindowCluster());
   g_bdOnDevice = true;
   g_bd_regions++;
   g_bd_events += ev.size();
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Row N2: the two quadratic loops of BDData::UpdateBD (src/bddata.cpp:646-659).  ModifyRP and Summarize are free functions
 * of bddata.cpp, so the same interposition serves them: the sort stays here, with the reference's own comparator (the
 * order std::sort leaves among equal keys is then the reference's), the loops run on the device.
 * ------------------------------------------------------------------------------------------------------------------- */
bool Compare2RP(const RP_READ &first, const RP_READ &second);   /* src/bddata.cpp:239-253 */
namespace {
void numberChromosomes(const std::vector<RP_READ> &v, std::vector<int32_t> &ca, std::vector<int32_t> &cb)
{
   std::map<std::string, int32_t> id;
   ca.resize(v.size());
   cb.resize(v.size());
   for (size_t i = 0; i < v.size(); i++) {
      ca[i] = id.insert(std::make_pair(v[i].ChrNameA, (int32_t)id.size())).first->second;
      cb[i] = id.insert(std::make_pair(v[i].ChrNameB, (int32_t)id.size())).first->second;
   }
}
void marshalRP(const std::vector<RP_READ> &v, std::vector<pg_rp> &in)
{
   in.resize(v.size());
   for (size_t i = 0; i < v.size(); i++) {
      in[i].pos_a = v[i].PosA; in[i].pos_b = v[i].PosB; in[i].insert_size = v[i].InsertSize;
      in[i].read_length = v[i].ReadLength; in[i].da = (uint8_t)v[i].DA; in[i].db = (uint8_t)v[i].DB;
   }
}
} // namespace

void ModifyRP(std::vector<RP_READ> &Reads_RP)
{
   std::cout << "Reads_RP.size(): " << Reads_RP.size() << std::endl;
   if (Reads_RP.size() == 0) return;
   std::cout << "sorting read-pair" << std::endl;
   sort(Reads_RP.begin(), Reads_RP.end(), Compare2RP);
   std::cout << "sorting read-pair finished." << std::endl;
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   std::vector<pg_rp> in;
   marshalRP(Reads_RP, in);
   std::vector<pg_rp_cluster> out(in.size());
   { Tick tk(t_rp_lib); if (pg_rp_modify(g_ctx, (int32_t)in.size(), &in[0], &out[0]) != PG_OK) die("pg_rp_modify"); }
   for (size_t i = 0; i < Reads_RP.size(); i++) {
      RP_READ &r = Reads_RP[i];
      r.PosA = out[i].pos_a; r.PosA1 = out[i].pos_a1; r.PosB = out[i].pos_b; r.PosB1 = out[i].pos_b1;
      if (out[i].visited) r.Visited = true;
   }
   g_rp_rows += (unsigned long)Reads_RP.size();
}

void Summarize(std::vector<RP_READ> &Reads_RP)
{
   if (Reads_RP.empty()) return;
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   std::vector<pg_rp> in;
   marshalRP(Reads_RP, in);
   std::vector<pg_rp_cluster> st(in.size());
   for (size_t i = 0; i < Reads_RP.size(); i++) {
      const RP_READ &r = Reads_RP[i];
      st[i].pos_a = r.PosA; st[i].pos_a1 = r.PosA1; st[i].pos_b = r.PosB; st[i].pos_b1 = r.PosB1;
      st[i].n_identical = r.NumberOfIdentical; st[i].merged_into = (uint32_t)i;
      st[i].visited = r.Visited ? 1 : 0; st[i].report = r.Report ? 1 : 0; st[i].pad = 0;
   }
   { Tick tk(t_rp_lib); if (pg_rp_summarize(g_ctx, (int32_t)in.size(), &in[0], &st[0]) != PG_OK) die("pg_rp_summarize"); }
   for (size_t i = 0; i < Reads_RP.size(); i++) {
      RP_READ &r = Reads_RP[i];
      r.NumberOfIdentical = st[i].n_identical;
      r.Visited = st[i].visited != 0;
      r.Report = st[i].report != 0;
      if (st[i].merged_into != (uint32_t)i && r.Tags.size()) {      /* src/bddata.cpp:500-505: the absorbing row collects the Tags, in row order */
         RP_READ &first = Reads_RP[st[i].merged_into];
         first.Tags.insert(first.Tags.end(), r.Tags.begin(), r.Tags.end());
         r.Tags.clear();
      }
   }
}

/* the inter-chromosomal variants (src/bddata.cpp:439-468, 577-624; called at :737-738): same binding, chromosome names numbered */
void ModifyRP_InterChr(std::vector<RP_READ> &Reads_RP)
{
   std::cout << "Reads_RP.size(): " << Reads_RP.size() << std::endl;
   if (Reads_RP.size() == 0) return;
   std::cout << "sorting read-pair" << std::endl;
   sort(Reads_RP.begin(), Reads_RP.end(), Compare2RP);
   std::cout << "sorting InterChr read-pair finished." << std::endl;
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   std::vector<pg_rp> in;
   marshalRP(Reads_RP, in);
   std::vector<int32_t> ca, cb;
   numberChromosomes(Reads_RP, ca, cb);
   std::vector<pg_rp_cluster> out(in.size());
   if (pg_rp_modify_interchr(g_ctx, (int32_t)in.size(), &in[0], &ca[0], &cb[0], &out[0]) != PG_OK) die("pg_rp_modify_interchr");
   for (size_t i = 0; i < Reads_RP.size(); i++) {
      RP_READ &r = Reads_RP[i];
      r.PosA = out[i].pos_a; r.PosA1 = out[i].pos_a1; r.PosB = out[i].pos_b; r.PosB1 = out[i].pos_b1;
   }
   g_rp_rows += (unsigned long)Reads_RP.size();
}

void Summarize_InterChr(std::vector<RP_READ> &Reads_RP)
{
   if (Reads_RP.empty()) return;
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   std::vector<pg_rp> in;
   marshalRP(Reads_RP, in);
   std::vector<int32_t> ca, cb;
   numberChromosomes(Reads_RP, ca, cb);
   std::vector<pg_rp_cluster> st(in.size());
   for (size_t i = 0; i < Reads_RP.size(); i++) {
      const RP_READ &r = Reads_RP[i];
      st[i].pos_a = r.PosA; st[i].pos_a1 = r.PosA1; st[i].pos_b = r.PosB; st[i].pos_b1 = r.PosB1;
      st[i].n_identical = r.NumberOfIdentical; st[i].merged_into = (uint32_t)i;
      st[i].visited = r.Visited ? 1 : 0; st[i].report = r.Report ? 1 : 0; st[i].pad = 0;
   }
   if (pg_rp_summarize_interchr(g_ctx, (int32_t)in.size(), &in[0], &ca[0], &cb[0], &st[0]) != PG_OK) die("pg_rp_summarize_interchr");
   for (size_t i = 0; i < Reads_RP.size(); i++) {
      RP_READ &r = Reads_RP[i];
      r.NumberOfIdentical = st[i].n_identical;
      r.Visited = st[i].visited != 0;
      r.Report = st[i].report != 0;
      if (st[i].merged_into != (uint32_t)i) {          /* src/bddata.cpp:605-609: only when the absorbing row holds Tags */
         RP_READ &first = Reads_RP[st[i].merged_into];
         if (first.Tags.size()) {
            first.Tags.insert(first.Tags.end(), r.Tags.begin(), r.Tags.end());
            r.Tags.clear();
         }
      }
   }
}

/* ---------------------------------------------------------------------------------------------------------------------
 * SV drivers (row N1): pg_classify once per window, replay per driver.
 * ------------------------------------------------------------------------------------------------------------------- */
namespace {
struct SvWindowResults {
   const SPLIT_READ *base = NULL;   /* identity of the window the records belong to */
   size_t n = 0;
   unsigned windowStart = 0, windowEnd = 0;
   const Chromosome *chromosome = NULL;
   uint32_t mask = 0;
   std::vector<pg_sv_record> records;
   size_t first[8] = {0, 0, 0, 0, 0, 0, 0, 0};   /* records of driver d: [first[d], first[d+1]) */
} g_sv;

int chrIndexOf(const Chromosome *c)
{
   for (unsigned i = 0; g_genome.getChr(i) != NULL; i++) if (g_genome.getChr(i) == c) return (int)i;
   return -1;
}

void pushUps(std::vector<pg_up> &dst, const SortedUniquePoints &src)
{
   for (unsigned k = 0; k < src.size(); k++) {
      pg_up u;
      u.abs_loc = src[k].AbsLoc;
      u.len = src[k].LengthStr;
      u.mm = src[k].Mismatches;
      u.dir = src[k].Direction == BACKWARD ? 1 : 0;
      u.strand = src[k].Strand == ANTISENSE ? 1 : 0;
      u.chr = (int16_t)chrIndexOf(src[k].chromosome_p);
      dst.push_back(u);
   }
}

void classifyWindow(ControlState &cs, unsigned NumBoxes, const SearchWindow &window, uint32_t mask)
{
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   Tick tk_all(t_sv_total);
   std::vector<SPLIT_READ> &reads = cs.Reads_SR;
   const int n = (int)reads.size();
   pg_sv_window w;
   memset(&w, 0, sizeof w);
   w.driver_mask = mask;
   w.chr = chrIndexOf(window.getChromosome());
   w.window_end = window.getEnd();
   w.num_boxes = NumBoxes;
   w.box_size = BoxSize;
   const SearchRegion *region = userSettings->getRegion();
   w.region_start_defined = region->isStartDefined();
   w.region_start = region->isStartDefined() ? region->getStart() : 0;
   w.region_end_defined = region->isEndDefined();
   w.region_end = region->isEndDefined() ? region->getEnd() : 0;
   w.min_inversion_size = userSettings->MIN_IndelSize_Inversion;
   w.min_matched_bases = userSettings->Min_Num_Matched_Bases;
   g_sv.records.resize((size_t)n * 2 + 1024);
   pg_sv_results res;
   res.records = &g_sv.records[0];
   res.records_cap = (uint32_t)g_sv.records.size();
   res.n_records = 0;
   int rc = PG_ERR_STATE;
   const bool resident = n > 0 && g_farResidentBase == &reads[0] && g_farResidentN == (size_t)n && getenv("PG_BRIDGE_NO_RESIDENT") == NULL;
   std::vector<uint8_t> flags(n);
   for (int i = 0; i < n; i++)
      flags[i] = (uint8_t)((reads[i].Used ? PG_SVR_USED : 0u) | (reads[i].FragName != reads[i].FarFragName ? PG_SVR_OTHERCHR : 0u));
   if (resident) {
      /* the window's reads, their anchors and their UP_Far lists are still on the device from SearchFarEnds: send only the
         UP_Close lists (run-length coded) and the Used flags */
      std::vector<uint32_t> closeOff(n + 1, 0);
      std::vector<pg_run> closeRuns;
      std::vector<pg_up> ups;
      std::vector<uint8_t> used(n);
      for (int i = 0; i < n; i++) {
         ups.clear();
         pushUps(ups, reads[i].UP_Close);
         const uint32_t k = pg_compress_ups(ups.empty() ? NULL : &ups[0], (uint32_t)ups.size(), NULL);
         const size_t at = closeRuns.size();
         closeRuns.resize(at + k);
         if (k) pg_compress_ups(&ups[0], (uint32_t)ups.size(), &closeRuns[at]);
         closeOff[i + 1] = (uint32_t)closeRuns.size();
         used[i] = (uint8_t)(flags[i] & PG_SVR_USED);
      }
      for (int attempt = 0; attempt < 2; attempt++) {
         { Tick tk(t_sv_lib); rc = pg_classify_resident(g_ctx, &w, &used[0], &closeOff[0], closeRuns.empty() ? NULL : &closeRuns[0], &res); }
         if (rc != PG_ERR_RUNS_CAP) break;
         g_sv.records.resize(res.n_records);
         res.records = &g_sv.records[0];
         res.records_cap = res.n_records;
      }
      if (rc == PG_OK) g_sv_resident += (unsigned long)n;
   }
   if (rc == PG_ERR_STATE) {   /* nothing resident (other caller order, or a batch that went through the chunked path) */
      std::string blob;
      std::vector<uint32_t> off(n + 1, 0), runOff(n + 1, 0);
      std::vector<uint8_t> md(n);
      std::vector<int16_t> isz(n);
      std::vector<uint16_t> nClose(n);
      std::vector<pg_run> runs;
      std::vector<pg_up> ups;
      for (int i = 0; i < n; i++) {
         const SPLIT_READ &r = reads[i];
         blob += r.getUnmatchedSeq();
         off[i + 1] = (uint32_t)blob.size();
         md[i] = (uint8_t)r.MatchedD;
         isz[i] = r.InsertSize;
         for (int list = 0; list < 2; list++) {
            ups.clear();
            pushUps(ups, list == 0 ? r.UP_Close : r.UP_Far);
            const uint32_t k = pg_compress_ups(ups.empty() ? NULL : &ups[0], (uint32_t)ups.size(), NULL);
            const size_t at = runs.size();
            runs.resize(at + k);
            if (k) pg_compress_ups(&ups[0], (uint32_t)ups.size(), &runs[at]);
            if (list == 0) nClose[i] = (uint16_t)k;
         }
         runOff[i + 1] = (uint32_t)runs.size();
      }
      pg_sv_reads in;
      memset(&in, 0, sizeof in);
      in.n_reads = n;
      in.seq_blob = (const uint8_t *)blob.c_str();
      in.seq_off = &off[0];
      in.matched_d = n ? &md[0] : NULL;
      in.insert_size = n ? &isz[0] : NULL;
      in.rflags = n ? &flags[0] : NULL;
      in.run_off = &runOff[0];
      in.n_close_runs = n ? &nClose[0] : NULL;
      in.runs = runs.empty() ? NULL : &runs[0];
      for (int attempt = 0; attempt < 2; attempt++) {
         { Tick tk(t_sv_lib); rc = pg_classify(g_ctx, &in, &w, &res); }
         if (rc != PG_ERR_RUNS_CAP) break;
         g_sv.records.resize(res.n_records);
         res.records = &g_sv.records[0];
         res.records_cap = res.n_records;
      }
   }
   if (rc != PG_OK) die("pg_classify");
   g_sv.records.resize(res.n_records);
   size_t k = 0;
   for (int d = 0; d <= 7; d++) {
      while (k < g_sv.records.size() && (int)g_sv.records[k].driver < d) k++;
      g_sv.first[d] = k;
   }
   g_sv.base = n ? &reads[0] : NULL;
   g_sv.n = (size_t)n;
   g_sv.windowStart = window.getStart();
   g_sv.windowEnd = window.getEnd();
   g_sv.chromosome = window.getChromosome();
   g_sv.mask = mask;
   g_sv_reads += (unsigned long)n;
   g_sv_records += (unsigned long)res.n_records;
}

uint32_t svMaskFromSettings()
{
   uint32_t m = (1u << PG_SV_D) | (1u << PG_SV_DI) | (1u << PG_SV_SI);
   if (userSettings->Analyze_TD) m |= (1u << PG_SV_TD) | (1u << PG_SV_TD_NT);
   if (userSettings->Analyze_INV) m |= (1u << PG_SV_INV) | (1u << PG_SV_INV_NT);
   return m;
}

/* The records of driver d, applied the way the driver acts on its reads (src/search_variant.cpp:117-175 and the like). */
void replayDriver(int d, ControlState &cs, unsigned NumBoxes, const SearchWindow &window, std::vector<unsigned> Vars[], int counts[3])
{
   std::vector<SPLIT_READ> &reads = cs.Reads_SR;
   const bool sameWindow = g_sv.base == (reads.empty() ? NULL : &reads[0]) && g_sv.n == reads.size() &&
                           g_sv.windowStart == window.getStart() && g_sv.windowEnd == window.getEnd() &&
                           g_sv.chromosome == window.getChromosome() && ((g_sv.mask >> d) & 1u);
   if (d == PG_SV_D || !sameWindow) classifyWindow(cs, NumBoxes, window, d == PG_SV_D ? svMaskFromSettings() : (1u << d));
   for (size_t k = g_sv.first[d]; k < g_sv.first[d + 1]; k++) {
      const pg_sv_record &rec = g_sv.records[k];
      SPLIT_READ &read = reads[rec.read];
      if (rec.fields & PG_SVF_CORE) {
         read.Left = rec.left;
         read.Right = rec.right;
         read.BP = rec.bp;
         read.BPLeft = rec.bp_left;
         read.BPRight = rec.bp_right;
         read.IndelSize = rec.indel_size;
      }
      if (rec.fields & PG_SVF_NTSTR) {
         if (rec.nt_src == PG_NT_EMPTY || rec.nt_len == 0) read.NT_str = "";
         else {
            const std::string src = rec.nt_src == PG_NT_SEQ ? read.getUnmatchedSeq() : ReverseComplement(read.getUnmatchedSeq());
            const std::string s = src.substr(rec.nt_start, rec.nt_len);
            read.NT_str = s.substr(rec.nt_rot) + s.substr(0, rec.nt_rot);
         }
      }
      if (rec.fields & PG_SVF_NTSIZE) read.NT_size = rec.nt_size;
      if (rec.fields & PG_SVF_KILLED) read.Used = true;
      switch (rec.outcome) {
      case PG_SV_FUTURE:
         cs.FutureReads_SR.push_back(read);     /* saveReadForNextCycle, src/pindel.cpp:587-592 */
         read.Used = true;
         break;
      case PG_SV_BOX:
         Vars[rec.box].push_back(rec.read);
         read.Used = true;
         counts[0]++;
         counts[read.MatchedD == Plus ? 1 : 2]++;
         break;
      case PG_SV_BREAK:
         read.Used = true;
         break;
      default:
         break;
      }
   }
}
} // namespace

/* ---------------------------------------------------------------------------------------------------------------------
 * Row N4: every reporter orders the reads of a box with an exchange sort, marks duplicates and cuts the surviving reads into
 * events before it decides and prints (src/reporter.cpp:932-972 as functions for D / SI / TD, inlined in the other two).  All
 * five reporters are supplied below around ONE pg_box_sort_kind call each; bubblesortReads is bound only to count the boxes the
 * reference's own loops still order (PG_BRIDGE_HOST_SORT=1, or a box that mixes chromosomes).
 * ------------------------------------------------------------------------------------------------------------------- */
namespace {
unsigned long g_boxes_sorted = 0, g_boxes_host = 0;
} // namespace

void bubblesortReads(const std::vector<SPLIT_READ> &Reads, std::vector<unsigned> &VariantIndices)
{
   typedef void (*Fn)(const std::vector<SPLIT_READ> &, std::vector<unsigned> &);
   static Fn original = (Fn)dlsym(RTLD_NEXT, "_Z15bubblesortReadsRKSt6vectorI10SPLIT_READSaIS0_EERS_IjSaIjEE");
   if (!original) { std::fprintf(stderr, "pindel_gpu_demo: the reference's bubblesortReads was not found\n"); std::exit(EXIT_FAILURE); }
   g_boxes_host++;
   original(Reads, VariantIndices);
}

/* ---------------------------------------------------------------------------------------------------------------------
 * The DI and inversion reporters carry the ordering and the duplicate scan INLINED, with their own rules
 * (SortOutputDI, src/reporter.cpp:1709-1851; OutputSorter::DoSortAndOutputInversions, src/output_sorter.cpp:62-260), so
 * there is no function to bind: the two reporters are supplied here.  Ordering, duplicate marks and the cut of the
 * surviving reads into events come from ONE pg_box_sort_kind call for all boxes of the driver; what follows per event --
 * the inversion reporter's adjustment of the breakpoints to the largest member, the report / balance decisions, the text --
 * is the reference's per-event logic and its own output functions (ReportEvent, IsInversion, OutputDI, OutputShortInversion,
 * OutputSorter::ReportIndelEvents).  PG_BRIDGE_HOST_SORT=1: the reference's two functions run (dlsym(RTLD_NEXT)).
 * ------------------------------------------------------------------------------------------------------------------- */
#include "output_file_data.h"
#include "output_sorter.h"
extern OutputFileData deletionFileData;
bool IsInversion(const SPLIT_READ &read, const std::string &chromosome);                                     /* src/reporter.cpp:1698 */
bool IsGoodDeletion(std::vector<SPLIT_READ> &GoodIndels, Indel4output &OneIndelEvent, unsigned RealStart, unsigned RealEnd, ControlState &currentState);
void OutputDI(std::vector<SPLIT_READ> &DI, const std::string &TheInput, const unsigned int &C_S, const unsigned int &C_E, const unsigned int &RealStart,
              const unsigned int &RealEnd, std::ofstream &DeletionOutf);
void OutputShortInversion(std::vector<SPLIT_READ> &supportingReads, const std::string &chromosome, const unsigned int &indexOfFirstRead,
                          const unsigned int &indexOfLastRead, const unsigned int &RealStart, const unsigned int &RealEnd, std::ofstream &InversionOutF);
namespace {
unsigned long g_boxes_inlined = 0;      /* boxes of the DI / inversion reporters ordered on the device */
struct OrderedBox {
   unsigned box;
   std::vector<unsigned> reads;         /* the box's reads in their final order */
   std::vector<uint8_t> event;          /* PG_BOX_EVENT_* per position */
};
/* every box the reporter would process, through the device; the reporter's side effects on its inputs are applied here:
 * the box's new order and the UniqueRead marks (later drivers read them).  false (nothing touched): a box mixes chromosomes
 * (smaller() and the D reporter's event test compare FragName / FarFragName first) -- the reference's own function runs. */
bool orderBoxes(int kind, std::vector<SPLIT_READ> &Reads, std::vector<unsigned> Vars[], unsigned NumBoxes, std::vector<OrderedBox> &out)
{
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   std::vector<uint32_t> boxOff(1, 0);
   std::vector<pg_sort_key> keys;
   std::map<std::string, uint32_t> nameId;
   const bool byName = kind == PG_BOX_D || kind == PG_BOX_SI;      /* bubblesortReads + markDuplicates */
   for (unsigned b = 0; b < NumBoxes; b++) {
      const std::vector<unsigned> &v = Vars[b];
      if (v.size() < userSettings->NumRead2ReportCutOff || v.empty()) continue;
      for (size_t k = 0; k < v.size(); k++) {
         const SPLIT_READ &r = Reads[v[k]];
         if (byName && (r.FragName != Reads[v[0]].FragName || r.FarFragName != r.FragName)) { out.clear(); return false; }
         pg_sort_key key;
         memset(&key, 0, sizeof key);
         key.bp_left = r.BPLeft; key.bp_right = r.BPRight; key.indel_size = r.IndelSize; key.nt_size = r.NT_size; key.bp = r.BP;
         key.unique = r.UniqueRead ? 1 : 0;
         if (byName) {
            key.left = r.Left; key.right = r.Right;
            std::map<std::string, uint32_t>::iterator it = nameId.find(r.Name);
            if (it == nameId.end()) it = nameId.insert(std::make_pair(r.Name, (uint32_t)nameId.size())).first;
            key.name_id = it->second;
         } else {
            key.left = (int32_t)r.LeftMostPos; key.read_length = (uint16_t)r.getReadLength(); key.matched_d = (uint8_t)r.MatchedD;
         }
         keys.push_back(key);
      }
      OrderedBox ob;
      ob.box = b;
      out.push_back(ob);
      boxOff.push_back((uint32_t)keys.size());
   }
   if (out.empty()) return true;
   std::vector<uint32_t> order(keys.size());
   std::vector<uint8_t> non(keys.size()), event(keys.size());
   { Tick tk(t_box_lib); if (pg_box_sort_kind(g_ctx, kind, (uint32_t)out.size(), &boxOff[0], &keys[0], &order[0], &non[0], &event[0]) != PG_OK) die("pg_box_sort_kind"); }
   for (size_t i = 0; i < out.size(); i++) {
      std::vector<unsigned> &v = Vars[out[i].box];
      for (uint32_t k = boxOff[i]; k < boxOff[i + 1]; k++) {
         out[i].reads.push_back(v[order[k]]);
         out[i].event.push_back(event[k]);
         if (non[k]) Reads[v[order[k]]].UniqueRead = false;
      }
      v = out[i].reads;
      if (byName) g_boxes_sorted++; else g_boxes_inlined++;
   }
   return true;
}
/* GoodIndels and the events' index ranges in it, from the device's flags */
void goodReadsAndEvents(const std::vector<SPLIT_READ> &Reads, const OrderedBox &ob, std::vector<SPLIT_READ> &good, std::vector<std::pair<unsigned, unsigned> > &events)
{
   good.clear();
   events.clear();
   for (size_t k = 0; k < ob.reads.size(); k++) {
      if (!(ob.event[k] & PG_BOX_EVENT_GOOD)) continue;
      if (ob.event[k] & PG_BOX_EVENT_START) events.push_back(std::make_pair((unsigned)good.size(), (unsigned)good.size()));
      else events.back().second = (unsigned)good.size();
      good.push_back(Reads[ob.reads[k]]);
   }
}
} // namespace

/* The D, SI and TD reporters (src/reporter.cpp:1395-1570, 975-1100, 1157-1290) call bubblesortReads / markDuplicates per box and
 * then walk the unique reads to cut them into events.  Supplied here the same way as the two reporters below: one
 * pg_box_sort_kind call for all boxes gives order, marks and the event cut; per event the reference's own functions decide and
 * print.  (The two bound functions above stay for any other caller.) */
bool IsGoodTD(std::vector<SPLIT_READ> &GoodIndels, Indel4output &OneIndelEvent, unsigned RealStart, unsigned RealEnd, ControlState &currentState);
namespace {
/* OneIndelEvent as the D / TD walks leave it for the event [first, last] of GoodIndels (src/reporter.cpp:1446-1497, 1206-1246) */
Indel4output breakpointEvent(const std::vector<SPLIT_READ> &good, unsigned first, unsigned last, const std::string &CurrentChr)
{
   Indel4output ev;
   ev.ChrName = good[first].FragName;
   ev.Start = first; ev.End = last; ev.Support = last - first + 1;
   ev.BPLeft = good[first].BPLeft; ev.BPRight = good[first].BPRight;
   ev.RealStart = ev.BPLeft; ev.RealEnd = ev.BPRight;
   ev.WhetherReport = true;
   GetRealStart4Deletion(CurrentChr, ev.RealStart, ev.RealEnd);
   return ev;
}
} // namespace

void SortOutputD(ControlState &currentState, const unsigned &NumBoxes, const std::string &CurrentChr, std::vector<SPLIT_READ> &Reads, std::vector<unsigned> Deletions[],
                 std::ofstream &DeletionOutf)
{
   std::vector<OrderedBox> boxes;
   if (getenv("PG_BRIDGE_HOST_SORT") || !orderBoxes(PG_BOX_D, Reads, Deletions, NumBoxes, boxes)) {
      typedef void (*Fn)(ControlState &, const unsigned &, const std::string &, std::vector<SPLIT_READ> &, std::vector<unsigned>[], std::ofstream &);
      static Fn original = (Fn)dlsym(RTLD_NEXT, "_Z11SortOutputDR12ControlStateRKjRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEERSt6vectorI10SPLIT_READSaISC_EEPSB_IjSaIjEERSt14basic_ofstreamIcS6_E");
      if (!original) { std::fprintf(stderr, "pindel_gpu_demo: the reference's SortOutputD was not found\n"); std::exit(EXIT_FAILURE); }
      original(currentState, NumBoxes, CurrentChr, Reads, Deletions, DeletionOutf);
      return;
   }
   LOG_INFO(*logStream << "Sorting and outputing deletions ..." << std::endl);
   std::vector<SPLIT_READ> good;
   std::vector<std::pair<unsigned, unsigned> > events;
   for (size_t i = 0; i < boxes.size(); i++) {
      goodReadsAndEvents(Reads, boxes[i], good, events);
      for (size_t e = 0; e < events.size(); e++) {
         Indel4output ev = breakpointEvent(good, events[e].first, events[e].second, CurrentChr);
         if (ev.Support < userSettings->NumRead2ReportCutOff || !IsGoodDeletion(good, ev, ev.RealStart, ev.RealEnd, currentState)) continue;   /* :1539 */
         if (good[ev.Start].IndelSize < userSettings->BalanceCutoff || ReportEvent(good, ev.Start, ev.End)) {                                  /* :1543-1559 */
            OutputDeletions(good, CurrentChr, ev.Start, ev.End, ev.RealStart, ev.RealEnd, DeletionOutf);
            deletionFileData.increaseTemplateSvCounter();
         }
      }
   }
   LOG_INFO(*logStream << "Deletions: " << deletionFileData.getTemplateSvCounter() << std::endl << std::endl);
}

void SortOutputSI(ControlState &currentState, const unsigned &NumBoxes, const std::string &CurrentChr, std::vector<SPLIT_READ> &Reads, std::vector<unsigned> SIs[],
                  std::ofstream &SIsOutf)
{
   std::vector<OrderedBox> boxes;
   if (getenv("PG_BRIDGE_HOST_SORT") || !orderBoxes(PG_BOX_SI, Reads, SIs, NumBoxes, boxes)) {
      typedef void (*Fn)(ControlState &, const unsigned &, const std::string &, std::vector<SPLIT_READ> &, std::vector<unsigned>[], std::ofstream &);
      static Fn original = (Fn)dlsym(RTLD_NEXT, "_Z12SortOutputSIR12ControlStateRKjRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEERSt6vectorI10SPLIT_READSaISC_EEPSB_IjSaIjEERSt14basic_ofstreamIcS6_E");
      if (!original) { std::fprintf(stderr, "pindel_gpu_demo: the reference's SortOutputSI was not found\n"); std::exit(EXIT_FAILURE); }
      original(currentState, NumBoxes, CurrentChr, Reads, SIs, SIsOutf);
      return;
   }
   LOG_INFO(*logStream << "Sorting and outputing short insertions ..." << std::endl);
   std::vector<SPLIT_READ> good;
   std::vector<std::pair<unsigned, unsigned> > events;
   for (size_t i = 0; i < boxes.size(); i++) {
      goodReadsAndEvents(Reads, boxes[i], good, events);
      for (size_t e = 0; e < events.size(); e++) {
         Indel4output ev;
         ev.initialize(events[e].first, good[events[e].first]);       /* src/pindel.h:418-432 */
         ev.End = events[e].second;
         ev.complete();
         GetRealStart4Insertion(CurrentChr, ev.IndelStr, ev.RealStart, ev.RealEnd);
         if (ev.WhetherReport && ev.Support >= userSettings->NumRead2ReportCutOff && ev.RealStart < ev.RealEnd)                               /* :1074 */
            OutputSIs(good, CurrentChr, ev.Start, ev.End, ev.RealStart, ev.RealEnd, SIsOutf);
      }
   }
   LOG_INFO(*logStream << "Short insertions: " << NumberOfSIsInstances << std::endl << std::endl);
}

void SortAndOutputTandemDuplications(ControlState &currentState, const unsigned &NumBoxes, const std::string &CurrentChr, std::vector<SPLIT_READ> &AllReads,
                                     std::vector<unsigned> TDs[], std::ofstream &TDOutf, const bool nonTemplate)
{
   std::vector<OrderedBox> boxes;
   if (getenv("PG_BRIDGE_HOST_SORT") || !orderBoxes(PG_BOX_D, AllReads, TDs, NumBoxes, boxes)) {
      typedef void (*Fn)(ControlState &, const unsigned &, const std::string &, std::vector<SPLIT_READ> &, std::vector<unsigned>[], std::ofstream &, const bool);
      static Fn original = (Fn)dlsym(RTLD_NEXT, "_Z31SortAndOutputTandemDuplicationsR12ControlStateRKjRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEERSt6vectorI10SPLIT_READSaISC_EEPSB_IjSaIjEERSt14basic_ofstreamIcS6_Eb");
      if (!original) { std::fprintf(stderr, "pindel_gpu_demo: the reference's SortAndOutputTandemDuplications was not found\n"); std::exit(EXIT_FAILURE); }
      original(currentState, NumBoxes, CurrentChr, AllReads, TDs, TDOutf, nonTemplate);
      return;
   }
   if (nonTemplate) { LOG_INFO(*logStream << "Sorting and outputing tandem duplications with non-template sequence ..." << std::endl); }
   else { LOG_INFO(*logStream << "Sorting and outputing tandem duplications ..." << std::endl); }
   int countTandemDuplications = 0;
   std::vector<SPLIT_READ> good;
   std::vector<std::pair<unsigned, unsigned> > events;
   for (size_t i = 0; i < boxes.size(); i++) {
      goodReadsAndEvents(AllReads, boxes[i], good, events);
      for (size_t e = 0; e < events.size(); e++) {
         Indel4output ev = breakpointEvent(good, events[e].first, events[e].second, CurrentChr);
         if (ev.Support < userSettings->NumRead2ReportCutOff || !IsGoodTD(good, ev, ev.RealStart, ev.RealEnd, currentState)) continue;        /* :1262 */
         if (good[ev.Start].IndelSize < userSettings->BalanceCutoff || ReportEvent(good, ev.Start, ev.End)) {                                  /* :1263-1271 */
            OutputTDs(good, CurrentChr, ev.Start, ev.End, ev.RealStart, ev.RealEnd, TDOutf);
            NumberOfTDInstances++;
            countTandemDuplications++;
         }
      }
   }
   if (nonTemplate) { LOG_INFO(*logStream << "Tandem duplications with non-template sequence (TD_NT): " << countTandemDuplications << std::endl << std::endl); }
   else { LOG_INFO(*logStream << "Tandem duplications: " << NumberOfTDInstances << std::endl << std::endl); }
}

void SortOutputDI(ControlState &currentState, const unsigned &NumBoxes, const std::string &CurrentChr, std::vector<SPLIT_READ> &Reads, std::vector<unsigned> DI[],
                  std::ofstream &DIOutf, std::ofstream &InvOutf)
{
   if (getenv("PG_BRIDGE_HOST_SORT")) {
      typedef void (*Fn)(ControlState &, const unsigned &, const std::string &, std::vector<SPLIT_READ> &, std::vector<unsigned>[], std::ofstream &, std::ofstream &);
      static Fn original = (Fn)dlsym(RTLD_NEXT, "_Z12SortOutputDIR12ControlStateRKjRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEERSt6vectorI10SPLIT_READSaISC_EEPSB_IjSaIjEERSt14basic_ofstreamIcS6_ESL_");
      if (!original) { std::fprintf(stderr, "pindel_gpu_demo: the reference's SortOutputDI was not found\n"); std::exit(EXIT_FAILURE); }
      original(currentState, NumBoxes, CurrentChr, Reads, DI, DIOutf, InvOutf);
      return;
   }
   LOG_INFO(*logStream << "Sorting and outputing deletions with non-template sequences ..." << std::endl);
   std::cout << "Added: Sorting and outputing deletions with non-template sequences ..." << std::endl;      /* src/reporter.cpp:1715 */
   std::vector<OrderedBox> boxes;
   orderBoxes(PG_BOX_DI, Reads, DI, NumBoxes, boxes);
   std::vector<SPLIT_READ> good;
   std::vector<std::pair<unsigned, unsigned> > events;
   for (size_t i = 0; i < boxes.size(); i++) {
      goodReadsAndEvents(Reads, boxes[i], good, events);
      for (size_t e = 0; e < events.size(); e++) {
         const unsigned first = events[e].first, last = events[e].second;
         if (last - first + 1 < userSettings->NumRead2ReportCutOff) continue;                                /* :1821 */
         const SPLIT_READ &lead = good[first];
         if (!(lead.IndelSize < userSettings->BalanceCutoff || ReportEvent(good, first, last))) continue;   /* :1824-1825 */
         unsigned int realStart = lead.BPLeft, realEnd = lead.BPRight;
         if (IsInversion(lead, CurrentChr)) {
            OutputShortInversion(good, CurrentChr, first, last, realStart, realEnd, InvOutf);
         } else {
            Indel4output ev;             /* as the reference's walk leaves it (:1790-1812): BPRight stays the box's first */
            ev.ChrName = lead.FragName; ev.Start = first; ev.End = last; ev.IndelSize = lead.IndelSize; ev.NT_size = lead.NT_size;
            ev.BPLeft = lead.BPLeft; ev.BPRight = good[0].BPRight; ev.WhetherReport = true;
            if (IsGoodDeletion(good, ev, realStart, realEnd, currentState)) {
               OutputDI(good, CurrentChr, first, last, realStart, realEnd, DIOutf);
               deletionFileData.increaseNonTemplateSvCounter();
            }
         }
      }
   }
   LOG_INFO(*logStream << "deletions with non-template sequences: " << deletionFileData.getNonTemplateSvCounter() << std::endl << std::endl);
}

int OutputSorter::DoSortAndOutputInversions(ControlState &currentState, std::vector<SPLIT_READ> &Reads, std::vector<unsigned> Inv[], bool isNonTemplateInversion)
{
   if (getenv("PG_BRIDGE_HOST_SORT")) {
      typedef int (*Fn)(OutputSorter *, ControlState &, std::vector<SPLIT_READ> &, std::vector<unsigned>[], bool);
      static Fn original = (Fn)dlsym(RTLD_NEXT, "_ZN12OutputSorter25DoSortAndOutputInversionsER12ControlStateRSt6vectorI10SPLIT_READSaIS3_EEPS2_IjSaIjEEb");
      if (!original) { std::fprintf(stderr, "pindel_gpu_demo: the reference's DoSortAndOutputInversions was not found\n"); std::exit(EXIT_FAILURE); }
      return original(this, currentState, Reads, Inv, isNonTemplateInversion);
   }
   int reported = 0;
   std::vector<OrderedBox> boxes;
   orderBoxes(isNonTemplateInversion ? PG_BOX_INV_NT : PG_BOX_INV, Reads, Inv, NumBoxes, boxes);
   std::vector<SPLIT_READ> good;
   std::vector<std::pair<unsigned, unsigned> > events;
   std::vector<Indel4output> indelEvents;
   for (size_t i = 0; i < boxes.size(); i++) {
      goodReadsAndEvents(Reads, boxes[i], good, events);          /* every read of the box is "good" here (src/output_sorter.cpp:134-137) */
      indelEvents.clear();
      bool report = true;               /* OneIndelEvent.WhetherReport: set once per box, never set back (:148) */
      for (size_t e = 0; e < events.size(); e++) {
         const unsigned first = events[e].first, last = events[e].second;
         Indel4output ev;
         ev.Start = first; ev.End = last;
         ev.BPLeft = good[first].BPLeft; ev.BPRight = good[first].BPRight;       /* before the adjustment below */
         /* the members take the breakpoints of the largest one (:159-189, :209-236) */
         unsigned maxSize = 0;
         for (unsigned k = first; k <= last; k++) if (good[k].IndelSize > maxSize) maxSize = good[k].IndelSize;
         for (unsigned k = first; k <= last; k++) {
            SPLIT_READ &r = good[k];
            if (r.IndelSize / (float)maxSize < 0.95 || maxSize + 30 > r.getReadLength() + r.IndelSize) { report = false; break; }
            const short diff = (maxSize - r.IndelSize) / 2;
            r.IndelSize = maxSize;
            r.BPLeft = r.BPLeft - diff;
            r.BPRight = r.BPRight + diff;
            if (r.MatchedD == '+') { if (r.BP > diff) r.BP = r.BP - diff; }
            else if (r.BP + diff < r.getReadLengthMinus()) r.BP = r.BP + diff;
         }
         if (e + 1 < events.size()) { ev.RealStart = good[first].BPLeft; ev.RealEnd = good[first].BPRight; }   /* :190-191: after the adjustment */
         else { ev.RealStart = ev.BPLeft; ev.RealEnd = ev.BPRight; }                                          /* :237-238: the box's last event keeps its own */
         ev.Support = last - first + 1;
         ev.WhetherReport = report;
         if (report) indelEvents.push_back(ev);
      }
      if (!indelEvents.empty()) reported += ReportIndelEvents(currentState, indelEvents, good);
   }
   return reported;
}

int SearchVariant::Search(BDData &bdData, ControlState &currentState, const unsigned NumBoxes, const SearchWindow &window)
{
   std::vector<unsigned> *Vars = new std::vector<unsigned>[NumBoxes];
   LOG_INFO(*logStream << "Searching " << typeOfVariant << " ... " << std::endl);
   int counts[3] = {0, 0, 0};
   /* typeOfVariant cannot tell the two apart: the reference prints the base-class placeholder for both */
   replayDriver(dynamic_cast<SearchDeletions *>(this) != NULL ? PG_SV_D : PG_SV_SI, currentState, NumBoxes, window, Vars, counts);
   Count_Var += counts[0];
   Count_Var_Plus += counts[1];
   Count_Var_Minus += counts[2];
   LOG_INFO(*logStream << "Total: " << Count_Var << "\t+" << Count_Var_Plus << "\t-" << Count_Var_Minus << std::endl);
   outputResults(bdData, currentState, Vars, NumBoxes, window);
   delete[] Vars;
   return EXIT_SUCCESS;
}

int searchIndels(ControlState &currentState, unsigned NumBoxes, const SearchWindow &window)
{
   std::vector<unsigned> *DI = new std::vector<unsigned>[NumBoxes];
   LOG_INFO(*logStream << "Searching deletion-insertions ... " << std::endl);
   int counts[3] = {0, 0, 0};
   replayDriver(PG_SV_DI, currentState, NumBoxes, window, DI, counts);
   std::ofstream DeletionOutf(userSettings->getDOutputFilename().c_str(), std::ios::app);
   std::ofstream inversionsOutf(userSettings->getINVOutputFilename().c_str(), std::ios::app);
   SortOutputDI(currentState, NumBoxes, window.getChromosome()->getSeq(), currentState.Reads_SR, DI, DeletionOutf, inversionsOutf);
   DeletionOutf.close();
   delete[] DI;
   return EXIT_SUCCESS;
}

int searchTandemDuplications(ControlState &currentState, unsigned NumBoxes, const SearchWindow &window)
{
   std::vector<unsigned> *TD = new std::vector<unsigned>[NumBoxes];
   LOG_INFO(*logStream << "Searching tandem duplication events ... " << std::endl);
   int counts[3] = {0, 0, 0};
   replayDriver(PG_SV_TD, currentState, NumBoxes, window, TD, counts);
   std::ofstream TDOutf(userSettings->getTDOutputFilename().c_str(), std::ios::app);
   SortAndOutputTandemDuplications(currentState, NumBoxes, window.getChromosome()->getSeq(), currentState.Reads_SR, TD, TDOutf, false);
   delete[] TD;
   return EXIT_SUCCESS;
}

int searchTandemDuplicationsNT(ControlState &currentState, unsigned NumBoxes, const SearchWindow &window)
{
   std::vector<unsigned> *TD_NT = new std::vector<unsigned>[NumBoxes];
   LOG_INFO(*logStream << "Searching tandem duplication events with non-template sequence ... " << std::endl);
   int counts[3] = {0, 0, 0};
   replayDriver(PG_SV_TD_NT, currentState, NumBoxes, window, TD_NT, counts);
   std::ofstream TDOutf(userSettings->getTDOutputFilename().c_str(), std::ios::app);
   SortAndOutputTandemDuplications(currentState, NumBoxes, window.getChromosome()->getSeq(), currentState.Reads_SR, TD_NT, TDOutf, true);
   TDOutf.close();
   delete[] TD_NT;
   return EXIT_SUCCESS;
}

int searchInversions(ControlState &currentState, unsigned NumBoxes, const SearchWindow &window)
{
   std::vector<unsigned> *Inv = new std::vector<unsigned>[NumBoxes];
   LOG_INFO(*logStream << "Searching inversions ... " << std::endl);
   int counts[3] = {0, 0, 0};
   replayDriver(PG_SV_INV, currentState, NumBoxes, window, Inv, counts);
   std::ofstream InversionOutf(userSettings->getINVOutputFilename().c_str(), std::ios::app);
   SortOutputInv(currentState, NumBoxes, window.getChromosome()->getSeq(), currentState.Reads_SR, Inv, InversionOutf);
   delete[] Inv;
   return EXIT_SUCCESS;
}

int searchInversionsNT(ControlState &currentState, unsigned NumBoxes, const SearchWindow &window)
{
   std::vector<unsigned> *Inv_NT = new std::vector<unsigned>[NumBoxes];
   LOG_INFO(*logStream << "Searching inversions with non-template sequence ... " << std::endl);
   int counts[3] = {0, 0, 0};
   replayDriver(PG_SV_INV_NT, currentState, NumBoxes, window, Inv_NT, counts);
   std::ofstream InversionOutf(userSettings->getINVOutputFilename().c_str(), std::ios::app);
   SortOutputInv_NT(currentState, NumBoxes, window.getChromosome()->getSeq(), currentState.Reads_SR, Inv_NT, InversionOutf);
   delete[] Inv_NT;
   return EXIT_SUCCESS;
}

/* Row N3 (first half): the BAM reader behind the reference's htslib calls (oracle/shim/minihts.cpp) inflates the file
 * through pg_bgzf_inflate instead of zlib.  PG_BRIDGE_ZLIB=1 keeps zlib (the test compares both). */
extern "C" int (*minihts_inflate_hook)(const uint8_t *file, size_t len, std::vector<uint8_t> &out);
static unsigned long g_bgzf_bytes = 0, g_rpd_rows = 0;
static const std::vector<uint8_t> *g_lastStream = NULL;
static std::vector<pg_bam_rec> g_lastTable;
static int deviceInflate(const uint8_t *file, size_t len, std::vector<uint8_t> &out)
{
   std::lock_guard<std::mutex> lock(g_mutex);
   ensureContext();
   const int64_t total = pg_bgzf_inflated_size(file, len);
   if (total < 0) { std::fprintf(stderr, "pindel_gpu_demo: not a BGZF file\n"); return 1; }
   out.resize((size_t)total);
   size_t got = 0;
   { Tick tk(t_inflate_lib); if (pg_bgzf_inflate(g_ctx, file, len, out.empty() ? NULL : &out[0], out.size(), &got, 1) != PG_OK) die("pg_bgzf_inflate"); }
   g_bgzf_bytes += (unsigned long)got;
   g_lastStream = &out;          /* the reader's copy of the stream: valid until its sam_close */
   g_lastTable.clear();
   return 0;
}

/* ... and builds its record table (the walk behind sam_itr_next's region test) with pg_bam_index on the stream that
 * inflate left on the device. */
extern "C" int (*minihts_index_hook)(size_t first_record, std::vector<uint8_t> &rows24);
static unsigned long g_bam_records = 0;
static int deviceIndex(size_t first_record, std::vector<uint8_t> &rows24)
{
   std::lock_guard<std::mutex> lock(g_mutex);
   Tick tk(t_index_lib);
   uint64_t n = 0, first = 0;
   int32_t nref = 0;
   int rc = pg_bam_index(g_ctx, NULL, 0, &n, &first, &nref);
   if (rc != PG_OK && rc != PG_ERR_RUNS_CAP) die("pg_bam_index");
   if (first != (uint64_t)first_record) { std::fprintf(stderr, "pindel_gpu_demo: BAM header size differs (%llu vs %zu)\n", (unsigned long long)first, first_record); return 1; }
   rows24.resize((size_t)n * sizeof(pg_bam_rec));
   if (n && pg_bam_index(g_ctx, (pg_bam_rec *)&rows24[0], n, &n, NULL, NULL) != PG_OK) die("pg_bam_index");
   g_bam_records += (unsigned long)n;
   g_lastTable.assign((const pg_bam_rec *)rows24.data(), (const pg_bam_rec *)rows24.data() + n);
   return 0;
}

/* Row N2, first half: ReadInBamReads_RP_Discovery (src/reader.cpp:418-480) is a free function of reader.cpp; bound by symbol
 * like the others.  The file is opened through the reference's own htslib calls (so the stream is inflated and indexed on
 * the device by the two hooks above), the region iterator + build_record_RP_Discovery run as pg_bam_rp_discovery on the
 * resident stream, and the rows come back with the index of their record for the fields only the host needs (ReadName,
 * reference names).  With PG_BRIDGE_ZLIB=1 the reference's own function runs (found with dlsym(RTLD_NEXT)). */
extern unsigned int g_NumReadScanned;
typedef bool (*RpDiscoveryFn)(const char *, const std::string &, const std::string *, std::vector<RP_READ> &, std::vector<RP_READ> &, int, std::string,
                              const SearchWindow &);
bool ReadInBamReads_RP_Discovery(const char *bam_path, const std::string &FragName, const std::string *CurrentChrSeq, std::vector<RP_READ> &LeftReads,
                                 std::vector<RP_READ> &LeftReads_InterChr, int InsertSize, std::string Tag, const SearchWindow &currentWindow)
{
   if (minihts_index_hook == NULL) {
      static RpDiscoveryFn original = (RpDiscoveryFn)dlsym(RTLD_NEXT, "_Z27ReadInBamReads_RP_DiscoveryPKcRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEPS7_RSt6vectorI7RP_READSaISB_EESE_iS6_RK12SearchWindow");
      if (!original) { std::fprintf(stderr, "pindel_gpu_demo: the reference's ReadInBamReads_RP_Discovery was not found\n"); std::exit(EXIT_FAILURE); }
      return original(bam_path, FragName, CurrentChrSeq, LeftReads, LeftReads_InterChr, InsertSize, Tag, currentWindow);
   }
   samFile *fp = sam_open(bam_path, "r");
   if (!fp) return false;
   hts_idx_t *idx = sam_index_load(fp, bam_path);
   bam_hdr_t *header = sam_hdr_read(fp);
   const int tid = bam_name2id(header, FragName.c_str());
   {
      std::lock_guard<std::mutex> lock(g_mutex);
      pg_rp_query q;
      q.tid = tid; q.beg = (int32_t)currentWindow.getStart(); q.end = (int32_t)currentWindow.getEnd(); q.insert_size = InsertSize;
      q.min_anchor_quality = userSettings->minimalAnchorQuality;
      uint64_t n = 0, ni = 0, scanned = 0;
      Tick tk(t_rpd_lib);
      int rc = pg_bam_rp_discovery(g_ctx, &q, NULL, NULL, 0, &n, NULL, NULL, 0, &ni, &scanned);
      std::vector<pg_rp> rows(n ? n : 1), irows(ni ? ni : 1);
      std::vector<uint32_t> src(n ? n : 1), isrc(ni ? ni : 1);
      if (rc == PG_ERR_RUNS_CAP) rc = pg_bam_rp_discovery(g_ctx, &q, &rows[0], &src[0], n, &n, &irows[0], &isrc[0], ni, &ni, &scanned);
      if (rc != PG_OK) die("pg_bam_rp_discovery");
      g_NumReadScanned += (unsigned int)scanned;
      const uint8_t *stream = g_lastStream->data();
      for (int list = 0; list < 2; list++) {
         const uint64_t cnt = list ? ni : n;
         for (uint64_t i = 0; i < cnt; i++) {
            const pg_rp &r = list ? irows[i] : rows[i];
            const pg_bam_rec &br = g_lastTable[list ? isrc[i] : src[i]];
            const uint8_t *rec = stream + br.off + 4;
            const int32_t mtid = (int32_t)(rec[20] | rec[21] << 8 | rec[22] << 16 | (uint32_t)rec[23] << 24);
            RP_READ t;
            t.DA = (char)r.da; t.DB = (char)r.db;
            t.PosA = t.OriginalPosA = r.pos_a;
            t.PosB = t.OriginalPosB = r.pos_b;
            t.Experimental_InsertSize = InsertSize;
            t.ReadName = (const char *)(rec + 32);
            t.MQA = t.MQB = br.mapq;
            t.ChrNameA = header->target_name[br.tid];
            t.ChrNameB = header->target_name[mtid];
            t.InsertSize = InsertSize;
            t.Tag = Tag;
            t.Tags.clear();
            t.Tags.push_back(Tag);
            t.ReadLength = r.read_length;
            (list ? LeftReads_InterChr : LeftReads).push_back(t);
         }
      }
      g_rpd_rows += (unsigned long)(n + ni);
   }
   bam_hdr_destroy(header);
   hts_idx_destroy(idx);
   sam_close(fp);
   (void)CurrentChrSeq;
   return true;
}

/* Rest of row N3: ReadInBamReads_SR (src/reader.cpp:483-560), bound by symbol like the discovery above.  The region iterator,
 * the pairing of mates on their names (fetch_func_SR), the three predicates and build_record_SR / build_record_RefRead run as
 * pg_bam_sr_candidates on the resident stream; the host adds what only it needs (Name, read_group, Tag, FragName) from the
 * records the rows point at, keeps the reader's counters, and hands every read to the caller's ReadBuffer exactly as
 * build_record_SR does.  By default the candidates are also searched where the filter left them (pg_bam_sr_close_end:
 * GetCloseEnd on the device without the sequences travelling to the host and back), so the reads reach the ReadBuffer with
 * their close ends and its flush only post-processes them; PG_BRIDGE_SR_UNCHAINED=1 keeps the two steps apart (filter, then
 * the buffer's batched search).  With PG_BRIDGE_HOST_SR=1 or PG_BRIDGE_ZLIB=1 the reference's own function runs (found with
 * dlsym(RTLD_NEXT)). */
extern unsigned int g_NumReadInWindow, g_InWinPlus, g_InWinMinus;
void showReadStats(const std::vector<SPLIT_READ> &Reads, const std::vector<SPLIT_READ> &OneEndMappedReads);
static unsigned long g_sr_candidates = 0, g_sr_searched = 0;
typedef bool (*SrReaderFn)(const char *, const std::string &, const std::string *, std::vector<SPLIT_READ> &, std::vector<SPLIT_READ> &,
                           std::vector<REF_READ> &, int, std::string, const SearchWindow &, ReadBuffer &, bool);
static void auxString(const uint8_t *rec, uint32_t rec_len, const char *tag, std::string &out)      /* bam_aux_get + bam_aux2Z */
{
   const uint32_t l_name = rec[8], n_cigar = rec[12] | rec[13] << 8, l_seq = rec[16] | rec[17] << 8 | rec[18] << 16 | (uint32_t)rec[19] << 24;
   const uint8_t *p = rec + 32 + l_name + 4ul * n_cigar + (l_seq + 1) / 2 + l_seq, *end = rec + rec_len;
   while (p + 3 <= end) {
      const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
      p += 3;
      size_t n;
      if (ty == 'A' || ty == 'c' || ty == 'C') n = 1;
      else if (ty == 's' || ty == 'S') n = 2;
      else if (ty == 'i' || ty == 'I' || ty == 'f') n = 4;
      else if (ty == 'd') n = 8;
      else if (ty == 'Z' || ty == 'H') { const uint8_t *e = p; while (e < end && *e) e++; if (e >= end) return; n = (size_t)(e - p) + 1; }
      else if (ty == 'B') {
         if (p + 5 > end) return;
         const uint8_t sub = p[0];
         const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
         if (!w) return;
         n = 5 + w * (p[1] | p[2] << 8 | p[3] << 16 | (size_t)p[4] << 24);
      } else return;
      if (n > (size_t)(end - p)) return;
      if (t0 == (uint8_t)tag[0] && t1 == (uint8_t)tag[1]) {
         if (ty == 'Z') out = (const char *)p;
         return;
      }
      p += n;
   }
}
bool ReadInBamReads_SR(const char *bam_path, const std::string &FragName, const std::string *CurrentChrSeq, std::vector<SPLIT_READ> &LeftReads,
                       std::vector<SPLIT_READ> &OneEndMappedReads, std::vector<REF_READ> &RefSupportingReads, int InsertSize, std::string Tag,
                       const SearchWindow &window, ReadBuffer &readBuffer, bool verbose)
{
   static const bool host_reader = getenv("PG_BRIDGE_HOST_SR") != NULL;
   if (minihts_index_hook == NULL || host_reader) {
      static SrReaderFn original = (SrReaderFn)dlsym(RTLD_NEXT, "_Z17ReadInBamReads_SRPKcRKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEPS7_RSt6vectorI10SPLIT_READSaISB_EESE_RSA_I8REF_READSaISF_EEiS6_RK12SearchWindowR10ReadBufferb");
      if (!original) { std::fprintf(stderr, "pindel_gpu_demo: the reference's ReadInBamReads_SR was not found\n"); std::exit(EXIT_FAILURE); }
      return original(bam_path, FragName, CurrentChrSeq, LeftReads, OneEndMappedReads, RefSupportingReads, InsertSize, Tag, window, readBuffer, verbose);
   }
   samFile *fp = sam_open(bam_path, "r");
   if (!fp) return false;
   hts_idx_t *idx = sam_index_load(fp, bam_path);
   bam_hdr_t *header = sam_hdr_read(fp);
   const int tid = bam_name2id(header, FragName.c_str());
   std::vector<pg_sr_cand> cands;
   std::vector<pg_sr_ref> refs;
   std::vector<char> seqs;
   std::vector<pg_read_out> sr_out;        /* GetCloseEnd of the candidates, searched on the device where the filter left them */
   std::vector<pg_run> sr_runs;
   bool searched = false;
   uint32_t stride = 0;
   {
      std::lock_guard<std::mutex> lock(g_mutex);
      ensureContext();
      pg_sr_query q;
      memset(&q, 0, sizeof q);
      q.tid = tid; q.beg = (int32_t)window.getStart(); q.end = (int32_t)window.getEnd(); q.insert_size = InsertSize;
      q.min_anchor_quality = userSettings->minimalAnchorQuality;
      q.chr_len = (uint32_t)(CurrentChrSeq->size() - 2 * g_SpacerBeforeAfter);
      q.nm_limit = userSettings->NM;
      q.max_mismatch_rate = userSettings->MaximumAllowedMismatchRate;
      uint64_t n = 0, m = 0, scanned = 0;
      int rc = PG_ERR_UNSUPPORTED;
      Tick tk(t_sr_lib);
      if (getenv("PG_BRIDGE_SR_UNCHAINED") == NULL) {
         /* filter + GetCloseEnd in one call: the size-asking call runs the kernels, the second one only copies */
         pg_results res;
         memset(&res, 0, sizeof res);
         rc = pg_bam_sr_close_end(g_ctx, &q, g_genome.getChrID(FragName), NULL, 0, &n, NULL, 0, &stride, NULL, 0, &m, &scanned, &res);
         if (rc == PG_ERR_RUNS_CAP) {
            cands.resize(n ? n : 1); refs.resize(m ? m : 1); seqs.resize(n * stride + 1);
            sr_out.resize(n ? n : 1); sr_runs.resize(res.n_runs ? res.n_runs : 1);
            res.out = &sr_out[0]; res.runs = &sr_runs[0]; res.runs_cap = (uint32_t)sr_runs.size();
            rc = pg_bam_sr_close_end(g_ctx, &q, g_genome.getChrID(FragName), &cands[0], n, &n, &seqs[0], n * stride, &stride, &refs[0], m, &m, &scanned, &res);
         }
         searched = rc == PG_OK;
      }
      if (rc == PG_ERR_UNSUPPORTED) {     /* candidates the search kernel does not take (>= 500 bases): filter alone, the buffer searches */
         rc = pg_bam_sr_candidates(g_ctx, &q, NULL, 0, &n, NULL, 0, &stride, NULL, 0, &m, &scanned);
         if (rc == PG_ERR_RUNS_CAP) {
            cands.resize(n ? n : 1); refs.resize(m ? m : 1); seqs.resize(n * stride + 1);
            rc = pg_bam_sr_candidates(g_ctx, &q, &cands[0], n, &n, &seqs[0], n * stride, &stride, &refs[0], m, &m, &scanned);
         }
      }
      if (rc == PG_ERR_DATA) {            /* the reference prints its explanation and leaves (src/reader.cpp:866-874) */
         *logStream << "Error: the insert size is only " << InsertSize << " while a read is at least as long" << std::endl;
         exit(EXIT_FAILURE);
      }
      if (rc != PG_OK) die("pg_bam_sr_close_end / pg_bam_sr_candidates");
      cands.resize(n); refs.resize(m);
      g_NumReadScanned += (unsigned int)scanned;
      g_sr_candidates += (unsigned long)n;
      if (searched) g_sr_searched += (unsigned long)n;
   }
   Tick tk_host(t_sr_host);
   const uint8_t *stream = g_lastStream->data();
   /* the reference interleaves the two kinds of calls, but they fill separate containers: candidates go to the buffer in order ... */
   for (size_t k = 0; k < cands.size(); k++) {
      const pg_sr_cand &c = cands[k];
      const pg_bam_rec &ra = g_lastTable[c.anchor], &rr = g_lastTable[c.read];
      const uint8_t *arec = stream + ra.off + 4, *rrec = stream + rr.off + 4;
      const uint32_t alen = stream[ra.off] | stream[ra.off + 1] << 8 | stream[ra.off + 2] << 16 | (uint32_t)stream[ra.off + 3] << 24;
      SPLIT_READ t;
      t.MS = c.ms;
      t.Name = "@";
      t.Name.append((const char *)(rrec + 32));
      if (rr.flag & 0x40) t.Name.append("/1");
      else if (rr.flag & 0x80) t.Name.append("/2");
      auxString(arec, alen, "RG", t.read_group);
      t.setUnmatchedSeq(std::string(&seqs[k * (size_t)stride], c.read_length));
      t.MatchedRelPos = c.matched_rel_pos;
      t.MatchedD = (char)c.matched_d;
      t.InsertSize = InsertSize;
      t.Tag = Tag;
      t.FragName = header->target_name[ra.tid];
      g_NumReadInWindow++;
      if (t.MatchedD == Plus) g_InWinPlus++; else g_InWinMinus++;
      if (searched) {                     /* what GetCloseEnd leaves in the read (closeEndBatch above) */
         if (t.InsertSize > g_maxInsertSize) g_maxInsertSize = t.InsertSize;
         if (sr_out[k].flags & PG_READ_UNSUPPORTED) die("read not supported by the library");
         for (unsigned f = 0; f < PG_READ_NFLIPS(sr_out[k].flags); f++) t.setUnmatchedSeq(ReverseComplement(t.getUnmatchedSeq()));
         t.UP_Close.clear();
         appendRuns(t.UP_Close, &sr_runs[sr_out[k].run_off], sr_out[k].n_close_runs);
         if (!t.UP_Close.empty()) t.Used = false;
         g_close_calls++;
      }
      g_flushPresearched = searched;      /* the buffer may flush inside addRead */
      readBuffer.addRead(t);
   }
   g_flushPresearched = searched;
   /* ... and the reference-supporting reads to their vector */
   for (size_t k = 0; k < refs.size(); k++) {
      REF_READ r;
      r.Pos = refs[k].pos;
      r.Tag = Tag;
      r.MQ = refs[k].mq;
      r.FragName = header->target_name[tid];
      r.ReadLength = refs[k].read_length;
      RefSupportingReads.push_back(r);
   }
   readBuffer.flush();
   g_flushPresearched = false;
   if (verbose) showReadStats(LeftReads, OneEndMappedReads);
   bam_hdr_destroy(header);
   hts_idx_destroy(idx);
   sam_close(fp);
   return true;
}

static void reportAtExit()
{
   const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t_start).count();
   const double lib = t_close_lib.s + t_far_lib.s + t_sv_lib.s + t_box_lib.s + t_rp_lib.s + t_inflate_lib.s + t_index_lib.s + t_rpd_lib.s + t_sr_lib.s + t_bd_lib.s;
   const double host = t_close_marshal.s + t_close_apply.s + t_far_marshal.s + t_far_apply.s + (t_sv_total.s - t_sv_lib.s) + t_sr_host.s;
   std::fprintf(stderr, "pindel_gpu_demo timing: wall %.3f s; context + reference upload %.3f; library calls %.3f (close end %.3f in %lu, far end %.3f in %lu, "
                "classification %.3f in %lu, box ordering %.3f in %lu, pair clustering %.3f in %lu, inflate %.3f in %lu, record index %.3f in %lu, pair discovery %.3f in %lu, "
                "split-read filter + search %.3f in %lu, BreakDancer regions %.3f in %lu); bridge marshalling %.3f (close end: in %.3f, out %.3f; far end: in %.3f, out %.3f; classification records and replay %.3f; "
                "split-read rows to SPLIT_READs %.3f); the reference's own host code (parsing, window loop, reporters) %.3f\n",
                wall, t_setup.s, lib, t_close_lib.s, t_close_lib.n, t_far_lib.s, t_far_lib.n, t_sv_lib.s, t_sv_lib.n, t_box_lib.s, t_box_lib.n, t_rp_lib.s, t_rp_lib.n,
                t_inflate_lib.s, t_inflate_lib.n, t_index_lib.s, t_index_lib.n, t_rpd_lib.s, t_rpd_lib.n, t_sr_lib.s, t_sr_lib.n, t_bd_lib.s, t_bd_lib.n, host, t_close_marshal.s, t_close_apply.s,
                t_far_marshal.s, t_far_apply.s, t_sv_total.s - t_sv_lib.s, t_sr_host.s, wall - t_setup.s - lib - host);
   std::fprintf(stderr, "pindel_gpu_demo: %lu GetCloseEnd calls and %lu far-end reads served by libpindel_pg (%llu kernel launches); "
                "%lu reads classified by the SV drivers on the device, %lu records (%llu launches), %lu of the reads without a second upload; %lu read pairs clustered on the device (%llu launches); %lu BAM bytes inflated and %lu records indexed on the device, %lu discordant pairs found there; %lu boxes ordered on the device (%llu launches), %lu by the reference, %lu boxes of the DI / inversion reporters ordered and cut into events there; %lu split-read candidates from the device filter, %lu of them searched where the filter left them; %lu BreakDancer regions (%lu events) loaded into the library, %llu cluster lookups there\n",
                g_close_calls, g_far_reads, g_ctx ? (unsigned long long)pg_kernel_launches(g_ctx) : 0ull, g_sv_reads, g_sv_records,
                g_ctx ? (unsigned long long)pg_classify_launches(g_ctx) : 0ull, g_sv_resident, g_rp_rows,
                g_ctx ? (unsigned long long)pg_rp_launches(g_ctx) : 0ull, g_bgzf_bytes, g_bam_records, g_rpd_rows, g_boxes_sorted,
                g_ctx ? (unsigned long long)pg_box_sort_launches(g_ctx) : 0ull, g_boxes_host, g_boxes_inlined, g_sr_candidates, g_sr_searched, g_bd_regions, g_bd_events, g_ctx ? (unsigned long long)pg_bd_lookups(g_ctx) : 0ull);
}

int main(int argc, char *argv[])
{
   std::atexit(reportAtExit);        /* the reference's main() leaves through exit() */
   if (getenv("PG_BRIDGE_ZLIB") == NULL) { minihts_inflate_hook = deviceInflate; minihts_index_hook = deviceIndex; }
   return pindel_reference_main(argc, argv);
}
