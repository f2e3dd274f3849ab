/* oracle/ref_hook.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * In-process hook around the UNMODIFIED reference objects (compiled from /root/reference/src
 * by oracle/Makefile, pindel.cpp with -Dmain=pindel_reference_main).  It lets tests and the
 * CPU-baseline leg of bench.py call the reference's own seam functions on batches of reads:
 *
 *     GetCloseEnd(chrSeq, read)              src/pindel.h:523   (src/pindel.cpp:2531)
 *     CleanUniquePoints(read.UP_Close)       src/pindel.cpp:2904 (as reader.cpp:272 / :1541 do)
 *     SearchFarEnds(chrSeq, reads, chr)      src/pindel.h:741   (src/pindel.cpp:1115)
 *
 * and read back every UniquePoint they produce.  Nothing here restates the algorithm; it only
 * marshals plain C arrays into SPLIT_READ objects and back.
 *
 * pgref_classify does the same for the seven SV drivers SearchSVs runs (src/pindel.cpp:1140-1164).  What a driver hands
 * to the reporter (its Vars[box] lists) is captured instead of reported: SearchDeletions / SearchShortInsertions through
 * their virtual outputResults, the five free-function drivers through the SortOutput* definitions below, which take
 * precedence over reporter.o's inside this one library (oracle/Makefile links it with --allow-multiple-definition,
 * ref_hook.o first).  libpindel_refcore.so and pindel_ref keep the reference's reporter.
 */
#include "pindel.h"
#include "control_state.h"
#include "bddata.h"
#include "reader.h"
#include "searcher.h"
#include "user_defined_settings.h"
#include "logstream.h"
#include "reporter.h"
#include "search_variant.h"
#include "searchdeletions.h"
#include "searchshortinsertions.h"
#include "search_deletions_nt.h"
#include "search_tandem_duplications.h"
#include "search_tandem_duplications_nt.h"
#include "search_inversions.h"
#include "search_inversions_nt.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <omp.h>
#include <string>
#include <vector>

extern void init(int argc, char *argv[], ControlState &currentState);
extern void CleanUniquePoints(SortedUniquePoints &Input_UP);
extern std::vector<unsigned int> g_maxMismatch;
extern BDData g_bdData;
extern UserDefinedSettings *userSettings;

namespace {
ControlState *g_state = NULL;
std::vector<std::string> g_argv_store;

int chrIdOf(const Chromosome *c)
{
   for (unsigned i = 0;; i++) {
      const Chromosome *k = g_genome.getChr(i);
      if (k == NULL) return -1;
      if (k == c) return (int)i;
   }
}

void pushUP(std::vector<int32_t> &out, const UniquePoint &up)
{
   out.push_back(up.LengthStr);
   out.push_back((int32_t)up.AbsLoc);
   out.push_back(up.Direction == FORWARD ? 0 : 1);
   out.push_back(up.Strand == SENSE ? 0 : 1);
   out.push_back(up.Mismatches);
   out.push_back(chrIdOf(up.chromosome_p));
}

// ---- capture of the drivers' box lists ----
std::vector<int32_t> *g_varsSink = NULL;
int g_varsCount = 0;
void captureVars(std::vector<unsigned> Vars[], unsigned NumBoxes)
{
   if (g_varsSink == NULL) return;
   for (unsigned b = 0; b < NumBoxes; b++) {
      for (size_t k = 0; k < Vars[b].size(); k++) {
         g_varsSink->push_back((int32_t)b);
         g_varsSink->push_back((int32_t)Vars[b][k]);
         g_varsCount++;
      }
   }
}
class CaptureDeletions : public SearchDeletions
{
protected:
   void outputResults(BDData &, ControlState &, std::vector<unsigned> Vars[], const unsigned NumBoxes, const SearchWindow &)
   {
      captureVars(Vars, NumBoxes);
   }
};
class CaptureShortInsertions : public SearchShortInsertions
{
protected:
   void outputResults(BDData &, ControlState &, std::vector<unsigned> Vars[], const unsigned NumBoxes, const SearchWindow &)
   {
      captureVars(Vars, NumBoxes);
   }
};

uint32_t fnv1a(const std::string &s)
{
   uint32_t h = 2166136261u;
   for (size_t i = 0; i < s.size(); i++) h = (h ^ (uint8_t)s[i]) * 16777619u;
   return h;
}
void putState(std::vector<int32_t> &out, const SPLIT_READ &r)
{
   uint32_t head = 0, tail = 0;
   const std::string &s = r.NT_str;
   for (size_t i = 0; i < 4 && i < s.size(); i++) head |= (uint32_t)(uint8_t)s[i] << (8 * i);
   for (size_t i = 0; i < 4 && i < s.size(); i++) tail |= (uint32_t)(uint8_t)s[s.size() - 1 - i] << (8 * i);
   out.push_back(r.Left);
   out.push_back(r.Right);
   out.push_back(r.BP);
   out.push_back((int32_t)r.BPLeft);
   out.push_back((int32_t)r.BPRight);
   out.push_back((int32_t)r.IndelSize);
   out.push_back(r.NT_size);
   out.push_back(r.Used ? 1 : 0);
   out.push_back((int32_t)s.size());
   out.push_back((int32_t)fnv1a(s));
   out.push_back((int32_t)head);
   out.push_back((int32_t)tail);
}
} // namespace

// Definitions that shadow reporter.o's inside libpindel_ref.so only (see the header comment).
void SortOutputDI(ControlState &, const unsigned &NumBoxes, const std::string &, std::vector<SPLIT_READ> &, std::vector<unsigned> DI[],
                  std::ofstream &, std::ofstream &)
{
   captureVars(DI, NumBoxes);
}
void SortAndOutputTandemDuplications(ControlState &, const unsigned &NumBoxes, const std::string &, std::vector<SPLIT_READ> &,
                                     std::vector<unsigned> TDs[], std::ofstream &, const bool)
{
   captureVars(TDs, NumBoxes);
}
void SortOutputInv(ControlState &, const unsigned &NumBoxes, const std::string &, std::vector<SPLIT_READ> &, std::vector<unsigned> Inv[],
                   std::ofstream &)
{
   captureVars(Inv, NumBoxes);
}
void SortOutputInv_NT(ControlState &, const unsigned &NumBoxes, const std::string &, std::vector<SPLIT_READ> &,
                      std::vector<unsigned> Inv[], std::ofstream &)
{
   captureVars(Inv, NumBoxes);
}

extern "C" {

/* argv: pindel command line WITHOUT argv[0]; must carry -f <fasta> -p <pindel text file> -o <prefix>. */
int pgref_init(int argc, const char **argv)
{
   if (g_state != NULL) return -1; // the reference keeps process-global state; one init per process
   userSettings = UserDefinedSettings::Instance();
   g_state = new ControlState();
   g_argv_store.clear();
   g_argv_store.push_back("pindel_ref_hook");
   for (int i = 0; i < argc; i++) g_argv_store.push_back(argv[i]);
   std::vector<char *> av;
   for (size_t i = 0; i < g_argv_store.size(); i++) av.push_back(const_cast<char *>(g_argv_store[i].c_str()));
   init((int)av.size(), &av[0], *g_state);
   return 0;
}

int pgref_n_chr(void)
{
   int n = 0;
   while (g_genome.getChr((unsigned)n) != NULL) n++;
   return n;
}
const char *pgref_chr_name(int i) { return g_genome.getChr((unsigned)i)->getName().c_str(); }
const char *pgref_chr_seq(int i) { return g_genome.getChr((unsigned)i)->getSeq().c_str(); }
uint32_t pgref_chr_len(int i) { return g_genome.getChr((unsigned)i)->getCompSize(); }
void pgref_max_mismatch(uint32_t *out, int n)
{
   for (int i = 0; i < n; i++) out[i] = (i < (int)g_maxMismatch.size()) ? g_maxMismatch[i] : 0;
}
int pgref_setting_int(const char *which)
{
   std::string w(which);
   if (w == "ADDITIONAL_MISMATCH") return userSettings->ADDITIONAL_MISMATCH;
   if (w == "Min_Perfect_Match_Around_BP") return userSettings->Min_Perfect_Match_Around_BP;
   if (w == "MaxRangeIndex") return userSettings->MaxRangeIndex;
   if (w == "minClose") return userSettings->minClose;
   if (w == "numThreads") return userSettings->numThreads;
   if (w == "MIN_IndelSize_Inversion") return userSettings->MIN_IndelSize_Inversion;
   if (w == "Min_Num_Matched_Bases") return userSettings->Min_Num_Matched_Bases;
   return -1;
}
double pgref_setting_double(const char *which)
{
   std::string w(which);
   if (w == "MaximumAllowedMismatchRate") return userSettings->MaximumAllowedMismatchRate;
   if (w == "Seq_Error_Rate") return userSettings->Seq_Error_Rate;
   if (w == "sensitivity") return userSettings->sensitivity;
   return -1.0;
}

/* Mirrors main()'s per-window g_bdData.loadRegion(currentWindow.makePindelCoordinateCopy())
 * (src/pindel.cpp:1831,1853); start/end are biological (unpadded) coordinates. */
int pgref_set_window(int chr, uint32_t start, uint32_t end)
{
   const Chromosome *c = g_genome.getChr((unsigned)chr);
   if (c == NULL) return -1;
   SearchWindow w(c, start, end);
   g_bdData.loadRegion(w.makePindelCoordinateCopy());
   return 0;
}

/* BAM-mode main() calls g_bdData.UpdateBD(currentState) before loadRegion (src/pindel.cpp:1841-1845); it is what
 * moves the external (-b) BreakDancer events into the active event list.  With no discordant pairs loaded this
 * activates exactly the -b events, which lets tests exercise the multi-region cluster pass of SearchFarEnd. */
void pgref_update_bd(void) { g_bdData.UpdateBD(*g_state); }

void pgref_free(void *p) { free(p); }

/* Runs the reference on n reads (all anchored on chromosome chr, as in one Pindel window).
 * mode bit0: close end (always), bit1: far end, bit2: timing only (no output stream).  threads = OpenMP threads.
 * Output: malloc'ed int32 stream, per read
 *   [flip, L_after, n_close_raw, n_close_clean, n_far, n_cluster_regions,
 *    close_raw UPs.., close_clean UPs.., far UPs.. (6 ints each: len, absLoc, dir, strand, mm, chr),
 *    cluster regions.. (3 ints each: chr, start, end)]
 * seconds[0] = wall time of the GetCloseEnd loop, seconds[1] = wall time of SearchFarEnds. */
int64_t pgref_search(int n, const char *seq_blob, const uint32_t *seq_off, const char *matched_d,
                     const uint32_t *rel_pos, const int16_t *insert_size, int chr, int mode, int threads,
                     int32_t **out_stream, double *seconds)
{
   const Chromosome *c = g_genome.getChr((unsigned)chr);
   if (c == NULL) return -1;
   const std::string &chrSeq = c->getSeq();
   omp_set_num_threads(threads > 0 ? threads : 1);

   std::vector<SPLIT_READ> reads((size_t)n);
   std::vector<std::string> original((size_t)n);
   for (int i = 0; i < n; i++) {
      SPLIT_READ &r = reads[i];
      r.Name = "@r";
      r.FragName = c->getName();
      original[i].assign(seq_blob + seq_off[i], seq_off[i + 1] - seq_off[i]);
      r.setUnmatchedSeq(original[i]);
      r.MatchedD = matched_d[i];
      r.MatchedRelPos = rel_pos[i];
      r.MS = 60;
      r.InsertSize = insert_size[i];
      r.Tag = "S";
   }
   std::vector<SortedUniquePoints> rawClose((size_t)n);

   std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
   #pragma omp parallel for schedule(dynamic, 64)
   for (int i = 0; i < n; i++) {
      GetCloseEnd(chrSeq, reads[i]);
   }
   std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
   for (int i = 0; i < n; i++) {
      rawClose[i] = reads[i].UP_Close;
      if (reads[i].UP_Close.size()) {
         reads[i].Used = false;
         reads[i].UniqueRead = true;
         CleanUniquePoints(reads[i].UP_Close);
      }
   }
   // the reference only forwards reads with a close end to SearchFarEnds (reader.cpp:262-288)
   std::vector<SPLIT_READ> kept;
   std::vector<int> keptIndex;
   for (int i = 0; i < n; i++) {
      if (reads[i].UP_Close.size()) {
         kept.push_back(reads[i]);
         keptIndex.push_back(i);
      }
   }
   std::chrono::steady_clock::time_point t2 = std::chrono::steady_clock::now(), t3 = t2;
   if (mode & 2) {
      std::ostream *savedLog = logStream;
      t2 = std::chrono::steady_clock::now();
      SearchFarEnds(chrSeq, kept, *c);
      t3 = std::chrono::steady_clock::now();
      logStream = savedLog;
   }
   if (seconds) {
      seconds[0] = std::chrono::duration<double>(t1 - t0).count();
      seconds[1] = std::chrono::duration<double>(t3 - t2).count();
   }

   std::vector<int32_t> out;
   if (mode & 4) { // timing only: no result stream
      int32_t *buf0 = (int32_t *)malloc(4);
      *out_stream = buf0;
      return 0;
   }
   out.reserve((size_t)n * 64);
   size_t k = 0;
   for (int i = 0; i < n; i++) {
      const SPLIT_READ *r = &reads[i];
      bool isKept = (k < keptIndex.size() && keptIndex[k] == i);
      if (isKept) r = &kept[k];
      out.push_back(r->getUnmatchedSeq() == original[i] ? 0 : 1);
      out.push_back(r->getReadLength());
      out.push_back((int32_t)rawClose[i].size());
      out.push_back((int32_t)r->UP_Close.size());
      out.push_back((int32_t)r->UP_Far.size());
      size_t nRegionsSlot = out.size();
      out.push_back(0);
      for (unsigned j = 0; j < rawClose[i].size(); j++) pushUP(out, rawClose[i][j]);
      for (unsigned j = 0; j < r->UP_Close.size(); j++) pushUP(out, r->UP_Close[j]);
      for (unsigned j = 0; j < r->UP_Far.size(); j++) pushUP(out, r->UP_Far[j]);
      if (isKept) {
         const SearchWindowCluster &cl = g_bdData.getCorrespondingSearchWindowCluster(*r);
         out[nRegionsSlot] = (int32_t)cl.size();
         for (unsigned j = 0; j < cl.size(); j++) {
            out.push_back(chrIdOf(cl[j].getChromosome()));
            out.push_back((int32_t)cl[j].getStart());
            out.push_back((int32_t)cl[j].getEnd());
         }
         k++;
      }
   }
   int32_t *buf = (int32_t *)malloc(out.size() * sizeof(int32_t) + 4);
   memcpy(buf, out.data(), out.size() * sizeof(int32_t));
   *out_stream = buf;
   return (int64_t)out.size();
}

/* Runs the reference's SV drivers (the ones in driver_mask, bit = position in SearchSVs' order: D, DI, TD, TD_NT, INV,
 * INV_NT, SI; src/pindel.cpp:1140-1164) over n reads of one window on chromosome chr.
 * up: [k,5] int32 rows (len, absLoc, dir 0/1, strand 0/1, mm); a read's first n_close[i] rows are UP_Close, the rest UP_Far.
 * region_mode 1: userSettings' region becomes [region_start, region_end] (stays so for the rest of the process).
 * Output stream: see svo_classify in oracle/sv_oracle.c (same format).  seconds[d] = wall time of driver d. */
int64_t pgref_classify(int n, const char *seq_blob, const uint32_t *seq_off, const char *matched_d, const int16_t *insert_size,
                       const uint8_t *rflags, const int32_t *up, const uint32_t *up_off, const uint32_t *n_close, int chr,
                       uint32_t win_start, uint32_t win_end, uint32_t num_boxes, uint32_t box_size, uint32_t driver_mask,
                       int region_mode, uint32_t region_start, uint32_t region_end, int32_t **out_stream, double *seconds)
{
   const Chromosome *c = g_genome.getChr((unsigned)chr);
   if (c == NULL) return -1;
   ControlState cs;
   cs.Reads_SR.resize((size_t)n);
   for (int i = 0; i < n; i++) {
      SPLIT_READ &r = cs.Reads_SR[i];
      r.Name = "@" + std::to_string(i);
      r.FragName = c->getName();
      r.setUnmatchedSeq(std::string(seq_blob + seq_off[i], seq_off[i + 1] - seq_off[i]));
      r.MatchedD = matched_d[i];
      r.MS = 60;
      r.InsertSize = insert_size[i];
      r.Tag = "S";
      r.Used = (rflags[i] & 1) != 0;
      const uint32_t a = up_off[i], b = up_off[i + 1];
      for (uint32_t k = a; k < b; k++) {
         UniquePoint p(c, (short)up[5 * k], (unsigned)up[5 * k + 1], up[5 * k + 2] ? BACKWARD : FORWARD,
                       up[5 * k + 3] ? ANTISENSE : SENSE, (short)up[5 * k + 4]);
         if (k - a < n_close[i]) r.UP_Close.push_back(p);
         else r.UP_Far.push_back(p);
      }
      r.FarFragName = (rflags[i] & 2) ? std::string("@other") : (r.UP_Far.empty() ? std::string("") : r.FragName);
   }
   if (region_mode == 1) const_cast<SearchRegion *>(userSettings->getRegion())->SetRegion(c->getName(), region_start, region_end);
   BoxSize = box_size;
   SearchWindow window(c, win_start, win_end);
   std::vector<int32_t> out;
   std::ostream *savedLog = logStream;
   for (int d = 0; d < 7; d++) {
      if (seconds) seconds[d] = 0.0;
      if (!((driver_mask >> d) & 1)) continue;
      std::vector<int32_t> vars;
      g_varsSink = &vars;
      g_varsCount = 0;
      const size_t futureBefore = cs.FutureReads_SR.size();
      std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
      switch (d) {
      case 0: { CaptureDeletions s; s.Search(g_bdData, cs, num_boxes, window); break; }
      case 1: searchIndels(cs, num_boxes, window); break;
      case 2: searchTandemDuplications(cs, num_boxes, window); break;
      case 3: searchTandemDuplicationsNT(cs, num_boxes, window); break;
      case 4: searchInversions(cs, num_boxes, window); break;
      case 5: searchInversionsNT(cs, num_boxes, window); break;
      case 6: { CaptureShortInsertions s; s.Search(g_bdData, cs, num_boxes, window); break; }
      }
      if (seconds) seconds[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      g_varsSink = NULL;
      out.push_back(d);
      out.push_back(g_varsCount);
      out.push_back((int32_t)(cs.FutureReads_SR.size() - futureBefore));
      out.insert(out.end(), vars.begin(), vars.end());
      for (size_t k = futureBefore; k < cs.FutureReads_SR.size(); k++) { // a saved read is a copy; its Name carries the index
         const SPLIT_READ &f = cs.FutureReads_SR[k];
         out.push_back((int32_t)atoi(f.Name.c_str() + 1));
         putState(out, f);
      }
      for (int i = 0; i < n; i++) putState(out, cs.Reads_SR[i]);
   }
   logStream = savedLog;
   int32_t *buf = (int32_t *)malloc(out.size() * sizeof(int32_t) + 4);
   memcpy(buf, out.data(), out.size() * sizeof(int32_t));
   *out_stream = buf;
   return (int64_t)out.size();
}

} // extern "C"

/* Row N2 of SURVEY.md section 8(f): the reference's own discordant-pair clustering, exactly as BDData::UpdateBD chains it
 * (src/bddata.cpp:654-659): sort by (PosA, PosB), ModifyRP (sort by original position descending, InitializeA1B1, the
 * O(n^2) interval-merging loop, + ReadLength), Summarize (identical-interval groups, Report for >= 5).  One thread: the
 * reference's OpenMP loop updates rows other threads are reading (src/bddata.cpp:401-417), so only the serial order is defined.
 * out: n rows of 14 int32 in the order ModifyRP's sort leaves: input index, OriginalPosA, OriginalPosB, DA, DB, ReadLength,
 * InsertSize, PosA, PosA1, PosB, PosB1, Visited, Report, NumberOfIdentical; then tags: for every row the number of Tags it
 * holds followed by the input indices they came from (Summarize moves Tags into the row that absorbs another). */
bool SortByFirstAndThenSecondCoordinate(const RP_READ &read1, const RP_READ &read2);
void ModifyRP(std::vector<RP_READ> &Reads_RP);
void Summarize(std::vector<RP_READ> &Reads_RP);

extern "C" int64_t pgref_rp_cluster(int n, const uint32_t *pos_a, const uint32_t *pos_b, const char *da, const char *db,
                                    const int16_t *read_length, const uint32_t *insert_size, int32_t **out_stream, double *seconds)
{
   std::vector<RP_READ> v((size_t)n);
   for (int i = 0; i < n; i++) {
      RP_READ &r = v[i];
      r.ReadName = "r" + std::to_string(i);
      r.ChrNameA = "c";
      r.ChrNameB = "c";
      r.DA = da[i];
      r.DB = db[i];
      r.PosA = r.OriginalPosA = pos_a[i];
      r.PosB = r.OriginalPosB = pos_b[i];
      r.ReadLength = read_length[i];
      r.InsertSize = insert_size[i];
      r.Experimental_InsertSize = (unsigned)i;   /* not read by the functions under test: carries the input index */
      r.Tag = std::to_string(i);
      r.Tags.push_back(r.Tag);
   }
   const int saved = omp_get_max_threads();
   omp_set_num_threads(1);
   auto t0 = std::chrono::steady_clock::now();
   std::sort(v.begin(), v.end(), SortByFirstAndThenSecondCoordinate);
   ModifyRP(v);
   auto t1 = std::chrono::steady_clock::now();
   Summarize(v);
   auto t2 = std::chrono::steady_clock::now();
   omp_set_num_threads(saved);
   if (seconds) {
      seconds[0] = std::chrono::duration<double>(t1 - t0).count();
      seconds[1] = std::chrono::duration<double>(t2 - t1).count();
   }
   std::vector<int32_t> out;
   for (int i = 0; i < n; i++) {
      const RP_READ &r = v[i];
      const int32_t row[14] = {(int32_t)r.Experimental_InsertSize, (int32_t)r.OriginalPosA, (int32_t)r.OriginalPosB, r.DA, r.DB, r.ReadLength,
                               (int32_t)r.InsertSize, (int32_t)r.PosA, (int32_t)r.PosA1, (int32_t)r.PosB, (int32_t)r.PosB1, r.Visited ? 1 : 0,
                               r.Report ? 1 : 0, (int32_t)r.NumberOfIdentical};
      out.insert(out.end(), row, row + 14);
   }
   for (int i = 0; i < n; i++) {
      out.push_back((int32_t)v[i].Tags.size());
      for (size_t k = 0; k < v[i].Tags.size(); k++) out.push_back((int32_t)std::stoi(v[i].Tags[k]));
   }
   int32_t *buf = (int32_t *)malloc(out.size() * sizeof(int32_t) + 4);
   memcpy(buf, out.data(), out.size() * sizeof(int32_t));
   *out_stream = buf;
   return (int64_t)out.size();
}

/* Row N2, first half: the reference's own discordant-pair discovery for one BAM file and window
 * (ReadInBamReads_RP_Discovery, src/reader.cpp:418-480 -> build_record_RP_Discovery, :982-1097), through the BAM reader of
 * oracle/shim/minihts.cpp.  Needs pgref_init (the window's chromosome comes from the loaded genome).
 * out: 9 int32 per row -- list (0 = Reads_RP_Discovery, 1 = ..._InterChr), PosA, PosB, DA, DB, ReadLength, InsertSize, MQA,
 * FNV-1a of ReadName -- in the order the reference pushed them; then 1 int32: g_NumReadScanned's increase. */
bool ReadInBamReads_RP_Discovery(const char *bam_path, const std::string &FragName, const std::string *CurrentChrSeq, std::vector<RP_READ> &LeftReads,
                                 std::vector<RP_READ> &LeftReads_InterChr, int InsertSize, std::string Tag, const SearchWindow &currentWindow);
extern unsigned int g_NumReadScanned;

extern "C" int64_t pgref_rp_discovery(const char *bam_path, int chr, uint32_t start, uint32_t end, int insert_size, int32_t **out_stream)
{
   const Chromosome *c = g_genome.getChr((unsigned)chr);
   if (c == NULL) return -1;
   SearchWindow w(c, start, end);
   std::vector<RP_READ> intra, inter;
   const unsigned before = g_NumReadScanned;
   if (!ReadInBamReads_RP_Discovery(bam_path, c->getName(), &c->getSeq(), intra, inter, insert_size, "TAG", w)) return -2;
   std::vector<int32_t> out;
   for (int list = 0; list < 2; list++) {
      const std::vector<RP_READ> &v = list ? inter : intra;
      for (size_t i = 0; i < v.size(); i++) {
         uint32_t h = 2166136261u;
         for (size_t k = 0; k < v[i].ReadName.size(); k++) h = (h ^ (uint8_t)v[i].ReadName[k]) * 16777619u;
         const int32_t row[9] = {list, (int32_t)v[i].PosA, (int32_t)v[i].PosB, v[i].DA, v[i].DB, v[i].ReadLength, (int32_t)v[i].InsertSize,
                                 v[i].MQA, (int32_t)h};
         out.insert(out.end(), row, row + 9);
      }
   }
   out.push_back((int32_t)(g_NumReadScanned - before));
   int32_t *buf = (int32_t *)malloc(out.size() * sizeof(int32_t) + 4);
   memcpy(buf, out.data(), out.size() * sizeof(int32_t));
   *out_stream = buf;
   return (int64_t)out.size();
}

/* Row N4: the reference's own bubblesortReads + markDuplicates (src/reporter.cpp:932-972) on reads built from plain arrays.
 * key: 9 int32 per read -- BPLeft, BPRight, IndelSize, NT_size, BP, Left, Right, name id (Name = "n<id>"), UniqueRead.
 * box_off [n_boxes + 1] into idx (read indices); idx is reordered in place per box, unique_out[r] = UniqueRead afterwards. */
void bubblesortReads(const std::vector<SPLIT_READ> &Reads, std::vector<unsigned> &VariantIndices);
void markDuplicates(std::vector<SPLIT_READ> &Reads, const std::vector<unsigned> &VariantIndices);

extern "C" int pgref_box_sort(int n_reads, const int32_t *key, int n_boxes, const uint32_t *box_off, uint32_t *idx, uint8_t *unique_out)
{
   std::vector<SPLIT_READ> reads((size_t)n_reads);
   for (int i = 0; i < n_reads; i++) {
      const int32_t *k = key + 9 * i;
      SPLIT_READ &r = reads[i];
      r.FragName = "c";
      r.BPLeft = (unsigned)k[0]; r.BPRight = (unsigned)k[1]; r.IndelSize = (unsigned)k[2]; r.NT_size = (unsigned short)k[3]; r.BP = (short)k[4];
      r.Left = k[5]; r.Right = k[6];
      r.Name = "n" + std::to_string(k[7]);
      r.UniqueRead = k[8] != 0;
   }
   for (int b = 0; b < n_boxes; b++) {
      std::vector<unsigned> v(idx + box_off[b], idx + box_off[b + 1]);
      if (v.empty()) continue;       /* the reporters only call with at least NumRead2ReportCutOff reads; size 0 would underflow */
      bubblesortReads(reads, v);
      markDuplicates(reads, v);
      for (size_t k = 0; k < v.size(); k++) idx[box_off[b] + k] = v[k];
   }
   for (int i = 0; i < n_reads; i++) unique_out[i] = reads[i].UniqueRead ? 1 : 0;
   return 0;
}

/* The inter-chromosomal variants exactly as BDData::UpdateBD chains them (src/bddata.cpp:736-739): ModifyRP_InterChr
 * (its own sort included), Summarize_InterChr.  chr_a / chr_b: numbers turned into names "c<k>".  Output as pgref_rp_cluster
 * (14 int32 per row + tags), with ChrNameA / ChrNameB numbers appended after the tags (2 int32 per row). */
void ModifyRP_InterChr(std::vector<RP_READ> &Reads_RP);
void Summarize_InterChr(std::vector<RP_READ> &Reads_RP);

extern "C" int64_t pgref_rp_cluster_interchr(int n, const uint32_t *pos_a, const uint32_t *pos_b, const char *da, const char *db,
                                             const int16_t *read_length, const uint32_t *insert_size, const int32_t *chr_a, const int32_t *chr_b,
                                             int32_t **out_stream)
{
   std::vector<RP_READ> v((size_t)n);
   for (int i = 0; i < n; i++) {
      RP_READ &r = v[i];
      r.ReadName = "r" + std::to_string(i);
      r.ChrNameA = "c" + std::to_string(chr_a[i]);
      r.ChrNameB = "c" + std::to_string(chr_b[i]);
      r.DA = da[i];
      r.DB = db[i];
      r.PosA = r.OriginalPosA = pos_a[i];
      r.PosB = r.OriginalPosB = pos_b[i];
      r.ReadLength = read_length[i];
      r.InsertSize = insert_size[i];
      r.Experimental_InsertSize = (unsigned)i;
      r.Tag = std::to_string(i);
      r.Tags.push_back(r.Tag);
   }
   ModifyRP_InterChr(v);
   Summarize_InterChr(v);
   std::vector<int32_t> out;
   for (int i = 0; i < n; i++) {
      const RP_READ &r = v[i];
      const int32_t row[14] = {(int32_t)r.Experimental_InsertSize, (int32_t)r.OriginalPosA, (int32_t)r.OriginalPosB, r.DA, r.DB, r.ReadLength,
                               (int32_t)r.InsertSize, (int32_t)r.PosA, (int32_t)r.PosA1, (int32_t)r.PosB, (int32_t)r.PosB1, r.Visited ? 1 : 0,
                               r.Report ? 1 : 0, (int32_t)r.NumberOfIdentical};
      out.insert(out.end(), row, row + 14);
   }
   for (int i = 0; i < n; i++) {
      out.push_back((int32_t)v[i].Tags.size());
      for (size_t k = 0; k < v[i].Tags.size(); k++) out.push_back((int32_t)std::stoi(v[i].Tags[k]));
   }
   for (int i = 0; i < n; i++) {
      out.push_back(std::stoi(v[i].ChrNameA.substr(1)));
      out.push_back(std::stoi(v[i].ChrNameB.substr(1)));
   }
   int32_t *buf = (int32_t *)malloc(out.size() * sizeof(int32_t) + 4);
   memcpy(buf, out.data(), out.size() * sizeof(int32_t));
   *out_stream = buf;
   return (int64_t)out.size();
}

/* Groundwork for the rest of row N3 (the reader's split-read filter): what ReadInBamReads_SR (src/reader.cpp:483-560) hands to
 * the ReadBuffer for one BAM file and window -- fetch_func_SR's mate pairing (:1099-1151), the predicates isGoodAnchor /
 * isWeirdRead / isRefRead (:561-690) and build_record_SR / build_record_RefRead (:799-923) -- captured before GetCloseEnd
 * sees it.  ReadBuffer::addRead is supplied here (this object comes first on the link line, see oracle/Makefile): same
 * behaviour as src/read_buffer.cpp:104-111, plus the capture.
 * out: number of candidates, per candidate 7 int32 (FNV-1a of Name, MatchedD, MatchedRelPos, MS, InsertSize, ReadLength, FNV-1a of
 * UnmatchedSeq), number of RefSupportingReads, per read 3 int32 (Pos, MQ, ReadLength). */
#include "read_buffer.h"
static std::vector<SPLIT_READ> *g_capture = NULL;
void ReadBuffer::addRead(const SPLIT_READ &newRead)
{
   if (g_capture) {
      g_capture->push_back(newRead);
      return;
   }
   if (m_currentsize == m_CAPACITY) flush();
   m_rawreads.push_back(newRead);
   m_currentsize++;
}
bool ReadInBamReads_SR(const char *bam_path, const std::string &FragName, const std::string *CurrentChrSeq, std::vector<SPLIT_READ> &LeftReads,
                       std::vector<SPLIT_READ> &OneEndMappedReads, std::vector<REF_READ> &RefSupportingReads, int InsertSize, std::string Tag,
                       const SearchWindow &window, ReadBuffer &readBuffer, bool verbose);
extern "C" int64_t pgref_sr_candidates(const char *bam_path, int chr, uint32_t start, uint32_t end, int insert_size, int min_anchor_quality,
                                       int32_t **out_stream)
{
   const Chromosome *c = g_genome.getChr((unsigned)chr);
   if (c == NULL) return -1;
   SearchWindow w(c, start, end);
   std::vector<SPLIT_READ> left, one_end, captured;
   std::vector<REF_READ> ref_reads;
   ReadBuffer buffer(50000, left, one_end, c->getSeq());
   g_capture = &captured;
   const unsigned saved_quality = userSettings->minimalAnchorQuality;      /* -A: the one setting the filter reads */
   if (min_anchor_quality >= 0) userSettings->minimalAnchorQuality = (unsigned)min_anchor_quality;
   const bool ok = ReadInBamReads_SR(bam_path, c->getName(), &c->getSeq(), left, one_end, ref_reads, insert_size, "TAG", w, buffer, false);
   userSettings->minimalAnchorQuality = saved_quality;
   g_capture = NULL;
   if (!ok) return -2;
   std::vector<int32_t> out;
   out.push_back((int32_t)captured.size());
   for (size_t i = 0; i < captured.size(); i++) {
      const SPLIT_READ &r = captured[i];
      const int32_t row[7] = {(int32_t)fnv1a(r.Name), r.MatchedD, (int32_t)r.MatchedRelPos, r.MS, r.InsertSize, r.getReadLength(), (int32_t)fnv1a(r.getUnmatchedSeq())};
      out.insert(out.end(), row, row + 7);
   }
   out.push_back((int32_t)ref_reads.size());
   for (size_t i = 0; i < ref_reads.size(); i++) {
      out.push_back((int32_t)ref_reads[i].Pos); out.push_back(ref_reads[i].MQ); out.push_back(ref_reads[i].ReadLength);
   }
   int32_t *buf = (int32_t *)malloc(out.size() * sizeof(int32_t) + 4);
   memcpy(buf, out.data(), out.size() * sizeof(int32_t));
   *out_stream = buf;
   return (int64_t)out.size();
}
