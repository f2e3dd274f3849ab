// tests/cpp/host_mirror_stages_test.cpp -- drives the stage mirrors of include/pindel_pg.hpp (ModifyRP / Summarize, SortBoxes,
// InflateBam / IndexBam) the way the reference's callers use the functions they mirror, on inputs written by the pytest
// that launches it.
//   host_mirror_stages_test <in.bin> <out.bin>
// in.bin : int32 n_rp, {uint32 PosA, PosB, InsertSize, int16 ReadLength, char DA, DB}*,
//          int32 n_reads, {uint32 BPLeft, BPRight, IndelSize, int32 Left, Right, uint32 name, uint16 NT_size, int16 BP, uint8 unique}*,
//          int32 n_boxes, {int32 size, uint32 read*}*, int64 bam_bytes, bytes
// out.bin: int32 stream: per RP row (in ModifyRP's order) [input index, PosA, PosA1, PosB, PosB1, Visited, Report, NumberOfIdentical, n Tags];
//          per box its reordered read indices; per read UniqueRead; int64 inflated size, int64 n records, per record off(lo32), pos, end;
//          then the inversion reporter's loops on the same boxes (SortBoxes(PG_BOX_INV_NT, ..), reads as they came in, with
//          LeftMostPos = Left, anchor strand and read length derived from the name number): per box the reordered indices,
//          per read UniqueRead, per box position the event flags
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/pindel_pg.hpp"

using namespace pindel_pg;

template <class T> static T rd(std::ifstream &f) { T v; f.read((char *)&v, sizeof v); return v; }

int main(int argc, char **argv)
{
   if (argc != 3) return 2;
   std::ifstream f(argv[1], std::ios::binary);
   const int32_t n_rp = rd<int32_t>(f);
   std::vector<RP_READ> rp((size_t)n_rp);
   for (int i = 0; i < n_rp; i++) {
      rp[i].PosA = rp[i].OriginalPosA = rd<uint32_t>(f);
      rp[i].PosB = rp[i].OriginalPosB = rd<uint32_t>(f);
      rp[i].InsertSize = rd<uint32_t>(f);
      rp[i].ReadLength = rd<int16_t>(f);
      rp[i].DA = rd<char>(f); rp[i].DB = rd<char>(f);
      rp[i].ReadName = std::to_string(i);            // carries the input index through the sort
      rp[i].Tags.push_back("t");
   }
   const int32_t n_reads = rd<int32_t>(f);
   std::vector<SPLIT_READ> reads((size_t)n_reads);
   for (int i = 0; i < n_reads; i++) {
      SPLIT_READ &r = reads[i];
      r.BPLeft = rd<uint32_t>(f); r.BPRight = rd<uint32_t>(f); r.IndelSize = rd<uint32_t>(f); r.Left = rd<int32_t>(f); r.Right = rd<int32_t>(f);
      r.Name = "n" + std::to_string(rd<uint32_t>(f));
      r.NT_size = rd<uint16_t>(f); r.BP = rd<int16_t>(f); r.UniqueRead = rd<uint8_t>(f) != 0;
      const unsigned name = (unsigned)std::stoul(r.Name.substr(1));
      r.LeftMostPos = (unsigned)r.Left;
      r.MatchedD = (name & 1u) ? '+' : '-';
      r.setUnmatchedSeq(std::string(name % 3u == 0 ? 101 : 100, 'A'));
   }
   const int32_t n_boxes = rd<int32_t>(f);
   std::vector<std::vector<unsigned> > boxes((size_t)n_boxes);
   for (int b = 0; b < n_boxes; b++) {
      const int32_t k = rd<int32_t>(f);
      boxes[b].resize((size_t)k);
      for (int j = 0; j < k; j++) boxes[b][j] = rd<uint32_t>(f);
   }
   const int64_t bam_bytes = rd<int64_t>(f);
   std::string bam((size_t)bam_bytes, '\0');
   f.read(&bam[0], bam_bytes);
   std::vector<SPLIT_READ> reads_inv(reads);
   std::vector<std::vector<unsigned> > boxes_inv;
   try {
      Searcher s;                                    // throws without a CUDA device (no CPU path)
      s.ModifyRP(rp);                                // src/bddata.cpp:657
      s.Summarize(rp);                               // src/bddata.cpp:659
      boxes_inv = boxes;
      s.SortBoxes(reads, boxes);                     // src/reporter.cpp:1425-1427, all boxes at once
      std::vector<std::vector<uint8_t> > events;
      s.SortBoxes(PG_BOX_INV_NT, reads_inv, boxes_inv, events);      // src/output_sorter.cpp:74-205
      std::string inflated;
      const size_t total = s.InflateBam(bam, &inflated);
      const std::vector<pg_bam_rec> recs = s.IndexBam();
      std::vector<int32_t> o;
      for (int i = 0; i < n_rp; i++) {
         const RP_READ &r = rp[i];
         const int32_t row[9] = {std::stoi(r.ReadName), (int32_t)r.PosA, (int32_t)r.PosA1, (int32_t)r.PosB, (int32_t)r.PosB1, r.Visited, r.Report,
                                 (int32_t)r.NumberOfIdentical, (int32_t)r.Tags.size()};
         o.insert(o.end(), row, row + 9);
      }
      for (int b = 0; b < n_boxes; b++) o.insert(o.end(), boxes[b].begin(), boxes[b].end());
      for (int i = 0; i < n_reads; i++) o.push_back(reads[i].UniqueRead);
      o.push_back((int32_t)total); o.push_back((int32_t)(inflated.size() == total)); o.push_back((int32_t)recs.size());
      for (size_t i = 0; i < recs.size(); i++) { o.push_back((int32_t)recs[i].off); o.push_back(recs[i].pos); o.push_back(recs[i].end); }
      for (int b = 0; b < n_boxes; b++) o.insert(o.end(), boxes_inv[b].begin(), boxes_inv[b].end());
      for (int i = 0; i < n_reads; i++) o.push_back(reads_inv[i].UniqueRead);
      for (int b = 0; b < n_boxes; b++) o.insert(o.end(), events[b].begin(), events[b].end());
      std::ofstream out(argv[2], std::ios::binary);
      out.write((const char *)o.data(), (std::streamsize)(o.size() * sizeof(int32_t)));
   } catch (const std::exception &e) {
      std::cerr << "host_mirror_stages_test: " << e.what() << std::endl;
      return 1;
   }
   return 0;
}
