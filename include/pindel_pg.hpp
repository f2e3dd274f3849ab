// include/pindel_pg.hpp -- header-only C++ host mirror of the reference's search seam on top of the C ABI
// (include/pindel_pg.h).  Names, argument meaning and side effects follow the reference (citations under
// /root/reference/): a maintainer who knows
//
//     void GetCloseEnd(const std::string& CurrentChrSeq, SPLIT_READ& read)                    src/pindel.h:523
//     void SearchFarEnds(const std::string&, std::vector<SPLIT_READ>&, const Chromosome&)     src/pindel.h:741
//     struct UniquePoint / class SortedUniquePoints / struct SPLIT_READ                       src/pindel.h:137-197,265-383
//     std::string ReverseComplement(const std::string&)                                       src/pindel.cpp:2037-2048
//     void CleanUniquePoints(SortedUniquePoints&)                                             src/pindel.cpp:2904-2941
//
// finds the same things here, batched: Searcher::GetCloseEnd(reads) and Searcher::SearchFarEnds(reads, clusters) take the
// whole vector.  The stages either side of the search follow the same rule (reference name, batched argument):
//
//     void ModifyRP(std::vector<RP_READ>&) / void Summarize(std::vector<RP_READ>&)            src/bddata.cpp:385,470
//     void bubblesortReads(Reads, VariantIndices) / void markDuplicates(Reads, VariantIndices) src/reporter.cpp:932,946
//     htslib's BGZF inflate + record walk behind sam_itr_next                                  src/reader.cpp:1258-1346
//
//     bool ReadInBamReads_SR(bam_path, FragName, ..., window, readBuffer, verbose)             src/reader.cpp:483-560
//
//     void BDData::loadRegion(const SearchWindow&) / getCorrespondingSearchWindowCluster(read)  src/bddata.cpp:852,949
//
// -> Searcher::loadRegion (then SearchFarEnds(reads) needs no clusters), Searcher::ModifyRP / Summarize (whole vector),
// Searcher::SortBoxes (all boxes of a driver; with a kind: any reporter's loops and its event cut), Searcher::InflateBam /
// IndexBam / ReadInBamReads_SR (one window of the resident stream).  No search logic lives in this header; without a CUDA
// device the Searcher constructor throws.
#ifndef PINDEL_PG_HPP
#define PINDEL_PG_HPP
#include <algorithm>
#include <cctype>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "pindel_pg.h"

namespace pindel_pg {

const char SENSE = '+', ANTISENSE = '-', FORWARD = '+', BACKWARD = '-';        // src/pindel.h:110-113
const unsigned g_SpacerBeforeAfter = 100000;                                     // src/pindel.h:122

struct UniquePoint {                     // src/pindel.h:137-158; the chromosome is an index into the reference list
   int chromosome_id;
   short LengthStr;
   unsigned int AbsLoc;
   char Direction;
   char Strand;
   short Mismatches;
};

class SortedUniquePoints {               // src/pindel.h:160-197
public:
   void push_back(const UniquePoint &up) { m_positions.push_back(up); }
   unsigned int size() const { return (unsigned int)m_positions.size(); }
   bool empty() const { return m_positions.empty(); }
   void clear() { m_positions.clear(); }
   unsigned int MaxLen() const { return m_positions.empty() ? 0u : (unsigned int)m_positions.back().LengthStr; }   // src/pindel.cpp:490-498
   const UniquePoint &operator[](unsigned int i) const { return m_positions[i]; }
   UniquePoint &operator[](unsigned int i) { return m_positions[i]; }
   void swap(SortedUniquePoints &o) { m_positions.swap(o.m_positions); }
private:
   std::vector<UniquePoint> m_positions;
};

inline std::string ReverseComplement(const std::string &in)     // src/pindel.cpp:2037-2048 with Convert2RC4N (:966-970)
{
   std::string out(in.size(), '\0');
   for (size_t j = 0; j < in.size(); j++) {
      char c = in[in.size() - 1 - j], r;
      switch (c) { case 'A': r = 'T'; break; case 'C': r = 'G'; break; case 'G': r = 'C'; break; case 'T': r = 'A'; break;
                   case 'N': r = 'N'; break; default: r = '\0'; }
      out[j] = r;
   }
   return out;
}

inline void CleanUniquePoints(SortedUniquePoints &ups)          // src/pindel.cpp:2904-2941
{
   if (ups.empty()) return;
   SortedUniquePoints kept;
   const UniquePoint last = ups[ups.size() - 1];
   const unsigned terminal = last.Direction == FORWARD ? last.AbsLoc - last.LengthStr : last.AbsLoc + last.LengthStr;
   for (unsigned i = 0; i < ups.size(); i++) {
      const UniquePoint &u = ups[i];
      if (u.chromosome_id != last.chromosome_id || u.Direction != last.Direction || u.Strand != last.Strand) continue;
      unsigned t = last.Direction == FORWARD ? u.AbsLoc - u.LengthStr : u.AbsLoc + u.LengthStr;
      if (t == terminal) kept.push_back(u);
   }
   ups.swap(kept);
}

struct SPLIT_READ {                      // the fields of src/pindel.h:265-383 that the search path reads or writes
   std::string Name;
   int chromosome_id = 0;                // FragName, as an index
   char MatchedD = 0;
   unsigned int MatchedRelPos = 0;
   short InsertSize = 0;
   short MS = 0;                         // the anchor's mapping quality (set by the reader)
   bool Used = false;
   bool MapperSplit = false;
   SortedUniquePoints UP_Close, UP_Far;
   // what the SV drivers assign and the reporters' per-box ordering reads (src/pindel.h:349-360)
   short BP = 0;
   int Left = 0, Right = 0;
   unsigned int BPLeft = 0, BPRight = 0, IndelSize = 0;
   unsigned short NT_size = 0;
   bool UniqueRead = true;
   unsigned int LeftMostPos = 0;          // read by the DI / inversion reporters' duplicate scans (src/reporter.cpp:1764-1772)

   void setUnmatchedSeq(const std::string &s)                    // src/pindel.cpp:142-169: trailing non-alphanumerics are stripped
   {
      UnmatchedSeq = s;
      while (!UnmatchedSeq.empty() && !std::isalnum((unsigned char)UnmatchedSeq[UnmatchedSeq.size() - 1])) UnmatchedSeq.resize(UnmatchedSeq.size() - 1);
   }
   const std::string &getUnmatchedSeq() const { return UnmatchedSeq; }
   short getReadLength() const { return (short)UnmatchedSeq.size(); }
   bool hasCloseEnd() const { return !UP_Close.empty(); }
   unsigned int getLastAbsLocCloseEnd() const { return UP_Close[UP_Close.size() - 1].AbsLoc; }        // src/pindel.cpp:474-477
   bool goodFarEndFound() const { return UP_Far.MaxLen() + UP_Close.MaxLen() >= UnmatchedSeq.size(); } // src/pindel.cpp:480-483
private:
   std::string UnmatchedSeq;
};

struct REF_READ {                        // src/pindel.h:199-214 without Tag / FragName (the caller knows them)
   unsigned Pos = 0;
   short MQ = 0;
   short ReadLength = 0;
};

struct RP_READ {                         // src/pindel.h:215-262: the fields the read-pair clustering reads or writes
   RP_READ() : DA('k'), DB('k'), PosA(0), PosB(0), OriginalPosA(0), OriginalPosB(0), PosA1(0), PosB1(0), InsertSize(0), ReadLength(0),
               NumberOfIdentical(0), Report(false), Visited(false) {}
   std::string ReadName;
   char DA, DB;
   unsigned PosA, PosB, OriginalPosA, OriginalPosB, PosA1, PosB1;
   std::vector<std::string> Tags;
   unsigned InsertSize;
   short ReadLength;
   unsigned NumberOfIdentical;
   bool Report, Visited;
};

inline bool Compare2RP(const RP_READ &first, const RP_READ &second)   // src/bddata.cpp:239-253
{
   if (first.OriginalPosA > second.OriginalPosA) return true;
   if (first.OriginalPosA == second.OriginalPosA) return first.OriginalPosB > second.OriginalPosB;
   return false;
}

struct SearchWindow { int chromosome_id; unsigned int start, end; };   // src/pindel.h:665-716, [start, end) padded coordinates
typedef std::vector<SearchWindow> SearchWindowCluster;
struct BreakDancerCoordinate { int chromosome_id; unsigned int position, position2; };   // src/control_state.h:40-68
struct BreakDancerEvent { BreakDancerCoordinate first, second; };                        // src/control_state.h:70-78

class Searcher {
public:
   explicit Searcher(const pg_settings &s) : ctx_(pg_create(&s))
   {
      if (!ctx_) throw std::runtime_error(std::string("pg_create: ") + pg_create_error());
   }
   Searcher() : ctx_(NULL)
   {
      pg_settings s;
      pg_default_settings(&s);
      ctx_ = pg_create(&s);
      if (!ctx_) throw std::runtime_error(std::string("pg_create: ") + pg_create_error());
   }
   ~Searcher() { pg_destroy(ctx_); }

   // the padded in-memory chromosome strings (Genome::loadAll, src/pindel.cpp:237-312), in chromosome_id order
   void setReference(const std::vector<std::string> &chromosomes)
   {
      std::vector<const char *> p;
      std::vector<uint32_t> l;
      for (size_t i = 0; i < chromosomes.size(); i++) { p.push_back(chromosomes[i].data()); l.push_back((uint32_t)chromosomes[i].size()); }
      check(pg_set_reference(ctx_, (int32_t)p.size(), p.data(), l.data()), "pg_set_reference");
   }

   std::vector<unsigned int> maxMismatchTable()                  // g_maxMismatch, src/pindel.cpp:794-819
   {
      std::vector<uint32_t> t(500);
      check(pg_max_mismatch_table(ctx_, t.data()), "pg_max_mismatch_table");
      return std::vector<unsigned int>(t.begin(), t.end());
   }

   // GetCloseEnd for every read (src/pindel.cpp:2531-2605): may reverse-complement the read, fills UP_Close (not cleaned).
   void GetCloseEnd(std::vector<SPLIT_READ> &reads)
   {
      Batch b(reads, NULL, false);
      run(pg_close_end, b, "pg_close_end");
      for (size_t i = 0; i < reads.size(); i++) {
         SPLIT_READ &r = reads[i];
         for (unsigned f = 0; f < PG_READ_NFLIPS(b.out[i].flags); f++) r.setUnmatchedSeq(ReverseComplement(r.getUnmatchedSeq()));
         r.UP_Close.clear();
         append(r.UP_Close, b.runs.data() + b.out[i].run_off, b.out[i].n_close_runs);
         if (!r.UP_Close.empty()) r.Used = false;               // src/pindel.cpp:2294, :2321
      }
   }

   // SearchFarEnds (src/pindel.cpp:1115-1138) for reads whose UP_Close is set (already cleaned, as in the reference);
   // clusters[i] = g_bdData.getCorrespondingSearchWindowCluster(reads[i]) or empty.
   void SearchFarEnds(std::vector<SPLIT_READ> &reads, const std::vector<SearchWindowCluster> *clusters = NULL)
   {
      Batch b(reads, clusters, true);
      run(pg_far_end, b, "pg_far_end");
      for (size_t i = 0; i < reads.size(); i++) {
         if (reads[i].MapperSplit || reads[i].UP_Close.empty()) continue;
         reads[i].UP_Far.clear();
         append(reads[i].UP_Far, b.runs.data() + b.out[i].run_off + b.out[i].n_close_runs, b.out[i].n_far_runs);
      }
   }

   uint64_t kernelLaunches() const { return pg_kernel_launches(ctx_); }

   // BDData::loadRegion (src/bddata.cpp:852-946): events = m_bdEvents as UpdateBD leaves it (sorted with sortOnFirstBDCoordinate:
   // chromosome NAME, position, then the second coordinate); chromosome_names[c] = name of chromosome c of the reference list.
   // From here on SearchFarEnds(reads) without clusters takes every read's cluster from this region on the device
   // (getCorrespondingSearchWindowCluster, :949-967) until unloadRegion().
   void loadRegion(const std::vector<BreakDancerEvent> &events, const std::vector<std::string> &chromosome_names, const SearchWindow &window)
   {
      std::vector<pg_bd_event> ev(events.size());
      for (size_t i = 0; i < events.size(); i++) {
         ev[i].chr_a = events[i].first.chromosome_id; ev[i].pos_a = events[i].first.position; ev[i].pos2_a = events[i].first.position2;
         ev[i].chr_b = events[i].second.chromosome_id; ev[i].pos_b = events[i].second.position; ev[i].pos2_b = events[i].second.position2;
      }
      std::vector<size_t> by_name(chromosome_names.size());
      for (size_t c = 0; c < by_name.size(); c++) by_name[c] = c;
      std::sort(by_name.begin(), by_name.end(), [&](size_t a, size_t b) { return chromosome_names[a] < chromosome_names[b]; });
      std::vector<int32_t> rank(chromosome_names.size());
      for (size_t k = 0; k < by_name.size(); k++) rank[by_name[k]] = (int32_t)k;
      check(pg_bd_load_region(ctx_, ev.empty() ? NULL : ev.data(), (uint32_t)ev.size(), rank.data(), window.chromosome_id, window.start, window.end),
            "pg_bd_load_region");
   }
   void unloadRegion() { check(pg_bd_unload(ctx_), "pg_bd_unload"); }
   // getCorrespondingSearchWindowCluster for a list of close-end positions (hosts that keep the clusters; tests)
   std::vector<SearchWindowCluster> getCorrespondingSearchWindowClusters(const std::vector<unsigned int> &lastAbsLocCloseEnd)
   {
      std::vector<uint32_t> off(lastAbsLocCloseEnd.size() + 1, 0);
      std::vector<pg_region> reg;
      uint64_t n = 0;
      int rc = pg_bd_clusters(ctx_, (uint32_t)lastAbsLocCloseEnd.size(), lastAbsLocCloseEnd.data(), off.data(), NULL, 0, &n);
      if (rc == PG_ERR_RUNS_CAP) {
         reg.resize((size_t)n);
         rc = pg_bd_clusters(ctx_, (uint32_t)lastAbsLocCloseEnd.size(), lastAbsLocCloseEnd.data(), off.data(), reg.data(), reg.size(), &n);
      }
      check(rc, "pg_bd_clusters");
      std::vector<SearchWindowCluster> out(lastAbsLocCloseEnd.size());
      for (size_t i = 0; i < out.size(); i++)
         for (uint32_t k = off[i]; k < off[i + 1]; k++) { SearchWindow w = {reg[k].chr, reg[k].start, reg[k].end}; out[i].push_back(w); }
      return out;
   }

   // ModifyRP (src/bddata.cpp:385-431): sort by original position (descending), InitializeA1B1, the interval-merging loop in
   // serial row order, + ReadLength, Visited for pairs closer than 500 -- all rows of one chromosome pair.
   void ModifyRP(std::vector<RP_READ> &Reads_RP)
   {
      if (Reads_RP.empty()) return;
      std::sort(Reads_RP.begin(), Reads_RP.end(), Compare2RP);
      std::vector<pg_rp> in;
      marshal(Reads_RP, in);
      std::vector<pg_rp_cluster> out(in.size());
      check(pg_rp_modify(ctx_, (int32_t)in.size(), in.data(), out.data()), "pg_rp_modify");
      for (size_t i = 0; i < Reads_RP.size(); i++) {
         RP_READ &r = Reads_RP[i];
         r.PosA = out[i].pos_a; r.PosA1 = out[i].pos_a1; r.PosB = out[i].pos_b; r.PosB1 = out[i].pos_b1;
         if (out[i].visited) r.Visited = true;
      }
   }

   // Summarize (src/bddata.cpp:470-575): identical events merged into their first row (NumberOfIdentical, Tags), Report for >= 5
   void Summarize(std::vector<RP_READ> &Reads_RP)
   {
      if (Reads_RP.empty()) return;
      std::vector<pg_rp> in;
      marshal(Reads_RP, in);
      std::vector<pg_rp_cluster> st(in.size());
      for (size_t i = 0; i < Reads_RP.size(); i++) {
         const RP_READ &r = Reads_RP[i];
         st[i].pos_a = r.PosA; st[i].pos_a1 = r.PosA1; st[i].pos_b = r.PosB; st[i].pos_b1 = r.PosB1; st[i].n_identical = r.NumberOfIdentical;
         st[i].merged_into = (uint32_t)i; st[i].visited = r.Visited; st[i].report = r.Report; st[i].pad = 0;
      }
      check(pg_rp_summarize(ctx_, (int32_t)in.size(), in.data(), st.data()), "pg_rp_summarize");
      for (size_t i = 0; i < Reads_RP.size(); i++) {
         RP_READ &r = Reads_RP[i];
         r.NumberOfIdentical = st[i].n_identical; r.Visited = st[i].visited != 0; r.Report = st[i].report != 0;
         if (st[i].merged_into != (uint32_t)i && !r.Tags.empty()) {
            RP_READ &first = Reads_RP[st[i].merged_into];
            first.Tags.insert(first.Tags.end(), r.Tags.begin(), r.Tags.end());
            r.Tags.clear();
         }
      }
   }

   // bubblesortReads + markDuplicates (src/reporter.cpp:932-972) for every box at once: boxes[b] is reordered and
   // Reads[..].UniqueRead cleared exactly as the two functions would do box by box.  Reads of one chromosome.
   void SortBoxes(std::vector<SPLIT_READ> &Reads, std::vector<std::vector<unsigned> > &boxes)
   {
      std::vector<uint32_t> off(1, 0);
      std::vector<pg_sort_key> keys;
      std::map<std::string, uint32_t> ids;
      for (size_t b = 0; b < boxes.size(); b++) {
         for (size_t k = 0; k < boxes[b].size(); k++) {
            const SPLIT_READ &r = Reads[boxes[b][k]];
            pg_sort_key key;
            std::memset(&key, 0, sizeof key);
            key.bp_left = r.BPLeft; key.bp_right = r.BPRight; key.indel_size = r.IndelSize; key.left = r.Left; key.right = r.Right;
            key.nt_size = r.NT_size; key.bp = r.BP; key.unique = r.UniqueRead ? 1 : 0;
            key.name_id = ids.insert(std::make_pair(r.Name, (uint32_t)ids.size())).first->second;
            keys.push_back(key);
         }
         off.push_back((uint32_t)keys.size());
      }
      if (keys.empty()) return;
      std::vector<uint32_t> order(keys.size());
      std::vector<uint8_t> non(keys.size());
      check(pg_box_sort(ctx_, (uint32_t)boxes.size(), off.data(), keys.data(), order.data(), non.data()), "pg_box_sort");
      for (size_t b = 0; b < boxes.size(); b++) {
         const std::vector<unsigned> before(boxes[b]);
         for (uint32_t k = off[b]; k < off[b + 1]; k++) {
            boxes[b][k - off[b]] = before[order[k]];
            if (non[k]) Reads[before[order[k]]].UniqueRead = false;
         }
      }
   }

   // The ordering loop, duplicate scan and event walk of ANY reporter (kind = PG_BOX_D / _SI / _DI / _INV / _INV_NT; the DI and
   // inversion reporters carry them inlined, src/reporter.cpp:1728-1813, src/output_sorter.cpp:74-205): boxes[b] reordered,
   // UniqueRead cleared, and events[b][k] = PG_BOX_EVENT_GOOD | PG_BOX_EVENT_START flags of the k-th read of box b afterwards.
   void SortBoxes(int kind, std::vector<SPLIT_READ> &Reads, std::vector<std::vector<unsigned> > &boxes, std::vector<std::vector<uint8_t> > &events)
   {
      std::vector<uint32_t> off(1, 0);
      std::vector<pg_sort_key> keys;
      std::map<std::string, uint32_t> ids;
      for (size_t b = 0; b < boxes.size(); b++) {
         for (size_t k = 0; k < boxes[b].size(); k++) {
            const SPLIT_READ &r = Reads[boxes[b][k]];
            pg_sort_key key;
            std::memset(&key, 0, sizeof key);
            key.bp_left = r.BPLeft; key.bp_right = r.BPRight; key.indel_size = r.IndelSize; key.nt_size = r.NT_size; key.bp = r.BP;
            key.unique = r.UniqueRead ? 1 : 0;
            if (kind == PG_BOX_D || kind == PG_BOX_SI) {
               key.left = r.Left; key.right = r.Right;
               key.name_id = ids.insert(std::make_pair(r.Name, (uint32_t)ids.size())).first->second;
            } else {
               key.left = (int32_t)r.LeftMostPos; key.read_length = (uint16_t)r.getReadLength(); key.matched_d = (uint8_t)r.MatchedD;
            }
            keys.push_back(key);
         }
         off.push_back((uint32_t)keys.size());
      }
      events.assign(boxes.size(), std::vector<uint8_t>());
      if (keys.empty()) return;
      std::vector<uint32_t> order(keys.size());
      std::vector<uint8_t> non(keys.size()), ev(keys.size());
      check(pg_box_sort_kind(ctx_, kind, (uint32_t)boxes.size(), off.data(), keys.data(), order.data(), non.data(), ev.data()), "pg_box_sort_kind");
      for (size_t b = 0; b < boxes.size(); b++) {
         const std::vector<unsigned> before(boxes[b]);
         for (uint32_t k = off[b]; k < off[b + 1]; k++) {
            boxes[b][k - off[b]] = before[order[k]];
            if (non[k]) Reads[before[order[k]]].UniqueRead = false;
            events[b].push_back(ev[k]);
         }
      }
   }

   // BGZF inflate of a whole BAM file on the device; the stream stays there for IndexBam (and comes back when wanted)
   size_t InflateBam(const std::string &file_bytes, std::string *inflated = NULL, bool check_crc = true)
   {
      size_t n = 0;
      if (inflated) {
         const int64_t total = pg_bgzf_inflated_size((const uint8_t *)file_bytes.data(), file_bytes.size());
         if (total < 0) throw std::runtime_error("pg_bgzf_inflated_size: not a BGZF file");
         inflated->resize((size_t)total);
      }
      check(pg_bgzf_inflate(ctx_, (const uint8_t *)file_bytes.data(), file_bytes.size(), inflated && !inflated->empty() ? (uint8_t *)&(*inflated)[0] : NULL,
                            inflated ? inflated->size() : 0, &n, check_crc ? 1 : 0), "pg_bgzf_inflate");
      return n;
   }

   // the record table (offset, refID, pos, end, flag, mapq) of the inflated stream: the walk + bam_endpos of an iterator
   std::vector<pg_bam_rec> IndexBam(uint64_t *first_record = NULL, int32_t *n_ref = NULL)
   {
      uint64_t n = 0;
      int rc = pg_bam_index(ctx_, NULL, 0, &n, first_record, n_ref);
      if (rc != PG_OK && rc != PG_ERR_RUNS_CAP) check(rc, "pg_bam_index");
      std::vector<pg_bam_rec> recs((size_t)n);
      if (n) check(pg_bam_index(ctx_, recs.data(), n, &n, first_record, n_ref), "pg_bam_index");
      return recs;
   }

   // the same for a run of BGZF blocks (one window's chunk): first_record = uoffset of the chunk's first record, n_ref =
   // bam_hdr_t::n_targets; *tail = where a record cut off by the end of the run starts (the stream's length if none)
   std::vector<pg_bam_rec> IndexBamRun(uint64_t first_record, int32_t n_ref, uint64_t *tail)
   {
      uint64_t n = 0, t = 0;
      int rc = pg_bam_index_run(ctx_, first_record, n_ref, NULL, 0, &n, &t);
      if (rc != PG_OK && rc != PG_ERR_RUNS_CAP) check(rc, "pg_bam_index_run");
      std::vector<pg_bam_rec> recs((size_t)n);
      if (n) check(pg_bam_index_run(ctx_, first_record, n_ref, recs.data(), n, &n, &t), "pg_bam_index_run");
      if (tail) *tail = t;
      return recs;
   }

   // ReadInBamReads_SR (src/reader.cpp:483-560) for one window of the stream InflateBam + IndexBam left on the device: the
   // reads build_record_SR would hand to the ReadBuffer are appended to Reads (Name, UnmatchedSeq, MatchedD, MatchedRelPos, MS,
   // InsertSize; chromosome_id = the caller's index of FragName), the reads build_record_RefRead keeps to RefSupportingReads.
   // inflated / recs: the host copies InflateBam / IndexBam returned (the names are read from them).  -> records scanned.
   uint64_t ReadInBamReads_SR(const std::string &inflated, const std::vector<pg_bam_rec> &recs, const pg_sr_query &q, int chromosome_id,
                              std::vector<SPLIT_READ> &Reads, std::vector<REF_READ> &RefSupportingReads)
   {
      uint64_t n = 0, m = 0, scanned = 0;
      uint32_t stride = 0;
      int rc = pg_bam_sr_candidates(ctx_, &q, NULL, 0, &n, NULL, 0, &stride, NULL, 0, &m, &scanned);
      if (rc != PG_OK && rc != PG_ERR_RUNS_CAP) check(rc, "pg_bam_sr_candidates");
      if (rc == PG_OK) return scanned;
      std::vector<pg_sr_cand> cands((size_t)n + 1);
      std::vector<pg_sr_ref> refs((size_t)m + 1);
      std::vector<char> seqs((size_t)n * stride + 1);
      check(pg_bam_sr_candidates(ctx_, &q, cands.data(), n, &n, seqs.data(), n * stride, &stride, refs.data(), m, &m, &scanned), "pg_bam_sr_candidates");
      for (uint64_t k = 0; k < n; k++) {
         const pg_sr_cand &c = cands[(size_t)k];
         const pg_bam_rec &rr = recs.at(c.read);
         SPLIT_READ t;
         t.Name = "@";
         t.Name.append(inflated.c_str() + rr.off + 36);
         if (rr.flag & 0x40) t.Name.append("/1");
         else if (rr.flag & 0x80) t.Name.append("/2");
         t.chromosome_id = chromosome_id;
         t.setUnmatchedSeq(std::string(&seqs[(size_t)k * stride], c.read_length));
         t.MatchedD = (char)c.matched_d;
         t.MatchedRelPos = c.matched_rel_pos;
         t.MS = c.ms;
         t.InsertSize = (short)q.insert_size;
         Reads.push_back(t);
      }
      for (uint64_t k = 0; k < m; k++) {
         REF_READ r;
         r.Pos = refs[(size_t)k].pos; r.MQ = refs[(size_t)k].mq; r.ReadLength = refs[(size_t)k].read_length;
         RefSupportingReads.push_back(r);
      }
      return scanned;
   }

private:
   static void marshal(const std::vector<RP_READ> &v, std::vector<pg_rp> &in)
   {
      in.resize(v.size());
      for (size_t i = 0; i < v.size(); i++) {
         in[i].pos_a = v[i].PosA; in[i].pos_b = v[i].PosB; in[i].insert_size = v[i].InsertSize; in[i].read_length = v[i].ReadLength;
         in[i].da = (uint8_t)v[i].DA; in[i].db = (uint8_t)v[i].DB;
      }
   }

   struct Batch {
      std::string blob;
      std::vector<uint32_t> off, pos, region_off;
      std::vector<char> md;
      std::vector<int16_t> isz;
      std::vector<int32_t> chr;
      std::vector<pg_close_info> ci;
      std::vector<pg_region> regions;
      std::vector<pg_read_out> out;
      std::vector<pg_run> runs;
      pg_reads in;
      Batch(const std::vector<SPLIT_READ> &reads, const std::vector<SearchWindowCluster> *clusters, bool far)
      {
         const size_t n = reads.size();
         off.assign(n + 1, 0); pos.resize(n); md.resize(n); isz.resize(n); chr.resize(n); ci.resize(n); region_off.assign(n + 1, 0);
         for (size_t i = 0; i < n; i++) {
            const SPLIT_READ &r = reads[i];
            blob += r.getUnmatchedSeq();
            off[i + 1] = (uint32_t)blob.size();
            pos[i] = r.MatchedRelPos; md[i] = r.MatchedD; isz[i] = r.InsertSize; chr[i] = r.chromosome_id;
            ci[i].pad = 0; ci[i].max_len = 0; ci[i].last_abs_loc = 0;
            if (far && !r.MapperSplit && !r.UP_Close.empty()) {
               ci[i].max_len = (uint16_t)r.UP_Close.MaxLen();
               ci[i].last_abs_loc = r.getLastAbsLocCloseEnd();
               if (clusters)
                  for (size_t k = 0; k < (*clusters)[i].size(); k++) {
                     pg_region g;
                     g.chr = (*clusters)[i][k].chromosome_id; g.start = (*clusters)[i][k].start; g.end = (*clusters)[i][k].end;
                     regions.push_back(g);
                  }
            }
            region_off[i + 1] = (uint32_t)regions.size();
         }
         std::memset(&in, 0, sizeof(in));
         in.n = (int32_t)n; in.seq_blob = blob.data(); in.seq_off = off.data(); in.chr_id = n ? chr.data() : NULL;
         if (!far) { in.matched_d = md.data(); in.matched_rel_pos = pos.data(); in.insert_size = isz.data(); }
         else { in.close_info = ci.data(); if (!regions.empty()) { in.regions = regions.data(); in.region_off = region_off.data(); } }
         out.resize(n ? n : 1);
         runs.resize(n * 8 + 1024);
      }
   };

   void run(int (*fn)(pg_ctx *, const pg_reads *, pg_results *), Batch &b, const char *what)
   {
      pg_results res;
      res.out = b.out.data(); res.runs = b.runs.data(); res.runs_cap = (uint32_t)b.runs.size(); res.n_runs = 0;
      int rc = fn(ctx_, &b.in, &res);
      if (rc == PG_ERR_RUNS_CAP) {
         b.runs.resize(res.n_runs);
         res.runs = b.runs.data(); res.runs_cap = (uint32_t)b.runs.size();
         rc = fn(ctx_, &b.in, &res);
      }
      check(rc, what);
      for (int i = 0; i < b.in.n; i++)
         if (b.out[i].flags & PG_READ_UNSUPPORTED) throw std::runtime_error(std::string(what) + ": read not supported by the library");
   }

   static void append(SortedUniquePoints &dst, const pg_run *runs, unsigned n)
   {
      for (unsigned r = 0; r < n; r++)
         for (unsigned j = 0; j < runs[r].count; j++) {
            UniquePoint u;
            u.chromosome_id = runs[r].chr;
            u.LengthStr = (short)(runs[r].len0 + j);
            u.AbsLoc = (runs[r].dirstrand & 1u) ? runs[r].abs_loc0 - j : runs[r].abs_loc0 + j;
            u.Direction = (runs[r].dirstrand & 1u) ? BACKWARD : FORWARD;
            u.Strand = (runs[r].dirstrand & 2u) ? ANTISENSE : SENSE;
            u.Mismatches = (short)runs[r].mm;
            dst.push_back(u);
         }
   }

   void check(int rc, const char *what)
   {
      if (rc != PG_OK) throw std::runtime_error(std::string(what) + ": " + pg_last_error(ctx_));
   }

   pg_ctx *ctx_;
   Searcher(const Searcher &);
   Searcher &operator=(const Searcher &);
};

} // namespace pindel_pg
#endif
