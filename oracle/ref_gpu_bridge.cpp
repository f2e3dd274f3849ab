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
 * GetCloseEnd is called once per read from inside the reference's own text-input loop (reader.cpp:248-256,300-306), so
 * for Pindel-text input this shim issues one-read batches (slow, but the caller stays untouched).  For BAM input the
 * reference funnels reads through ReadBuffer::flush (read_buffer.cpp:36-101), a public member function, so the bridge
 * supplies that too -- as the batched binding INTEGRATION.md describes: one pg_close_end call per flush of up to
 * 50 000 reads.  SearchFarEnds is a true batch call in both modes.
 */
#include "pindel.h"
#include "bddata.h"
#include "read_buffer.h"
#include "reader.h"
#include "logstream.h"
#include "user_defined_settings.h"

#include <cstdio>
#include <cstdlib>
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
std::mutex g_mutex;
unsigned long g_close_calls = 0, g_far_reads = 0;

void die(const char *what)
{
   std::fprintf(stderr, "pindel_gpu_demo: %s: %s\n", what, g_ctx ? pg_last_error(g_ctx) : pg_create_error());
   std::exit(EXIT_FAILURE);   /* the reference's own error convention: message + exit (e.g. src/pindel.cpp:674) */
}

void ensureContext()
{
   if (g_ctx) return;
   pg_settings s;
   pg_default_settings(&s);
   s.seq_error_rate = userSettings->Seq_Error_Rate;
   s.sensitivity = userSettings->sensitivity;
   s.max_mismatch_rate = userSettings->MaximumAllowedMismatchRate;
   s.additional_mismatch = userSettings->ADDITIONAL_MISMATCH;
   s.min_close = g_MinClose;
   s.min_perfect_match = userSettings->Min_Perfect_Match_Around_BP;
   s.max_range_index = userSettings->MaxRangeIndex;
   g_ctx = pg_create(&s);
   if (!g_ctx) die("pg_create");
   std::vector<const char *> seqs;
   std::vector<uint32_t> lens;
   for (unsigned i = 0; g_genome.getChr(i) != NULL; i++) {
      seqs.push_back(g_genome.getChr(i)->getSeq().c_str());
      lens.push_back(g_genome.getChr(i)->getCompSize());
   }
   if (pg_set_reference(g_ctx, (int)seqs.size(), &seqs[0], &lens[0]) != PG_OK) die("pg_set_reference");
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
   {
      std::lock_guard<std::mutex> lock(g_mutex);
      ensureContext();
      const int n = m_currentsize;
      std::string blob;
      std::vector<uint32_t> off(n + 1, 0), pos(n);
      std::vector<char> md(n);
      std::vector<int16_t> isz(n);
      std::vector<int32_t> chr(n);
      std::vector<int> idx;                  /* reads that take part (MapperSplit == false) */
      for (int i = 0; i < n; i++) {
         if (m_rawreads[i].MapperSplit) continue;
         const SPLIT_READ &r = m_rawreads[i];
         int k = (int)idx.size();
         blob += r.getUnmatchedSeq();
         off[k + 1] = (uint32_t)blob.size();
         md[k] = r.MatchedD; pos[k] = r.MatchedRelPos; isz[k] = r.InsertSize; chr[k] = g_genome.getChrID(r.FragName);
         if (r.InsertSize > g_maxInsertSize) g_maxInsertSize = r.InsertSize;
         idx.push_back(i);
      }
      const int m = (int)idx.size();
      if (m > 0) {
         pg_reads in;
         memset(&in, 0, sizeof(in));
         in.n = m; in.seq_blob = blob.c_str(); in.seq_off = &off[0]; in.matched_d = &md[0]; in.matched_rel_pos = &pos[0];
         in.insert_size = &isz[0]; in.chr_id = &chr[0];
         std::vector<pg_read_out> out(m);
         std::vector<pg_run> runs((size_t)m * 16 + 1024);
         pg_results res;
         res.out = &out[0]; res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size(); res.n_runs = 0;
         int rc = pg_close_end(g_ctx, &in, &res);
         if (rc == PG_ERR_RUNS_CAP) {
            runs.resize(res.n_runs);
            res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size();
            rc = pg_download(g_ctx, &res);
         }
         if (rc != PG_OK) die("pg_close_end (ReadBuffer::flush)");
         for (int k = 0; k < m; k++) {
            SPLIT_READ &read = m_rawreads[idx[k]];
            if (out[k].flags & PG_READ_UNSUPPORTED) die("read not supported by the library");
            for (unsigned f = 0; f < PG_READ_NFLIPS(out[k].flags); f++)
               read.setUnmatchedSeq(ReverseComplement(read.getUnmatchedSeq()));
            read.UP_Close.clear();
            appendRuns(read.UP_Close, &runs[out[k].run_off], out[k].n_close_runs);
            if (!read.UP_Close.empty()) read.Used = false;
         }
         g_close_calls += (unsigned long)m;
      }
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
   std::string blob;
   std::vector<uint32_t> off(n + 1, 0), region_off(n + 1, 0);
   std::vector<int32_t> chr(n);
   std::vector<pg_close_info> ci(n);
   std::vector<pg_region> regions;
   for (int i = 0; i < n; i++) {
      const SPLIT_READ &r = reads[i];
      blob += r.getUnmatchedSeq();
      off[i + 1] = (uint32_t)blob.size();
      chr[i] = g_genome.getChrID(r.FragName);
      ci[i].pad = 0;
      if (r.MapperSplit || r.UP_Close.empty()) { ci[i].max_len = 0; ci[i].last_abs_loc = 0; }   /* src/pindel.cpp:1129 */
      else {
         ci[i].max_len = (uint16_t)r.UP_Close.MaxLen();
         ci[i].last_abs_loc = r.getLastAbsLocCloseEnd();
         const SearchWindowCluster &cl = g_bdData.getCorrespondingSearchWindowCluster(r);      /* src/pindel.cpp:1006 */
         for (unsigned k = 0; k < cl.size(); k++) {
            pg_region g;
            g.chr = -1;
            for (unsigned c = 0; g_genome.getChr(c) != NULL; c++) if (g_genome.getChr(c) == cl[k].getChromosome()) g.chr = (int32_t)c;
            g.start = cl[k].getStart(); g.end = cl[k].getEnd();
            regions.push_back(g);
         }
      }
      region_off[i + 1] = (uint32_t)regions.size();
   }
   (void)currentChromosome;
   pg_reads in;
   memset(&in, 0, sizeof(in));
   in.n = n; in.seq_blob = blob.c_str(); in.seq_off = &off[0]; in.chr_id = n ? &chr[0] : NULL; in.close_info = n ? &ci[0] : NULL;
   if (!regions.empty()) { in.regions = &regions[0]; in.region_off = &region_off[0]; }
   std::vector<pg_read_out> out(n ? n : 1);
   std::vector<pg_run> runs((size_t)n * 16 + 1024);
   pg_results res;
   res.out = &out[0]; res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size(); res.n_runs = 0;
   int rc = pg_far_end(g_ctx, &in, &res);
   if (rc == PG_ERR_RUNS_CAP) {
      runs.resize(res.n_runs);
      res.runs = &runs[0]; res.runs_cap = (uint32_t)runs.size();
      rc = pg_download(g_ctx, &res);
   }
   if (rc != PG_OK) die("pg_far_end");
   for (int i = 0; i < n; i++) {
      if (ci[i].max_len == 0) continue;
      reads[i].UP_Far.clear();
      appendRuns(reads[i].UP_Far, &runs[out[i].run_off + out[i].n_close_runs], out[i].n_far_runs);
      g_far_reads++;
   }
   *logStream << "Far end searching completed for this window." << std::endl;   /* src/pindel.cpp:1137 */
}

static void reportAtExit()
{
   std::fprintf(stderr, "pindel_gpu_demo: %lu GetCloseEnd calls and %lu far-end reads served by libpindel_pg (%llu kernel launches)\n",
                g_close_calls, g_far_reads, g_ctx ? (unsigned long long)pg_kernel_launches(g_ctx) : 0ull);
}

int main(int argc, char *argv[])
{
   std::atexit(reportAtExit);        /* the reference's main() leaves through exit() */
   return pindel_reference_main(argc, argv);
}
