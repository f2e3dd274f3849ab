// tests/cpp/host_mirror_test.cpp -- drives include/pindel_pg.hpp the way the reference's callers drive the seam
// (ReadInRead: GetCloseEnd -> CleanUniquePoints -> keep reads with a close end; main: SearchFarEnds), on inputs written
// by the pytest that launches it, and dumps every UniquePoint in the layout of the oracle streams.
//   host_mirror_test <in.bin> <out.bin>
// in.bin : int32 n_chr, {uint32 len, bytes}*, int32 n_reads, {int32 len, bytes, char matched_d, uint32 pos, int16 isz, int32 chr}*
// out.bin: int32 stream [flip, L, n_close_raw, n_close_clean, n_far, 0, UPs.. (6 ints each)] per read
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/pindel_pg.hpp"

using namespace pindel_pg;

template <class T> static T rd(std::ifstream &f) { T v; f.read((char *)&v, sizeof v); return v; }

static void push(std::vector<int32_t> &o, const UniquePoint &u)
{
   o.push_back(u.LengthStr); o.push_back((int32_t)u.AbsLoc); o.push_back(u.Direction == FORWARD ? 0 : 1);
   o.push_back(u.Strand == SENSE ? 0 : 1); o.push_back(u.Mismatches); o.push_back(u.chromosome_id);
}

int main(int argc, char **argv)
{
   if (argc != 3) return 2;
   std::ifstream f(argv[1], std::ios::binary);
   int32_t n_chr = rd<int32_t>(f);
   std::vector<std::string> chroms((size_t)n_chr);
   for (int i = 0; i < n_chr; i++) { uint32_t l = rd<uint32_t>(f); chroms[i].resize(l); f.read(&chroms[i][0], l); }
   int32_t n = rd<int32_t>(f);
   std::vector<SPLIT_READ> reads((size_t)n);
   std::vector<std::string> original((size_t)n);
   for (int i = 0; i < n; i++) {
      int32_t l = rd<int32_t>(f);
      original[i].resize((size_t)l); f.read(&original[i][0], l);
      reads[i].setUnmatchedSeq(original[i]);
      reads[i].MatchedD = rd<char>(f); reads[i].MatchedRelPos = rd<uint32_t>(f); reads[i].InsertSize = rd<int16_t>(f);
      reads[i].chromosome_id = rd<int32_t>(f);
   }
   try {
      Searcher searcher;                        // Pindel defaults; throws without a CUDA device (no CPU path)
      searcher.setReference(chroms);
      searcher.GetCloseEnd(reads);              // src/reader.cpp:248-256
      std::vector<SortedUniquePoints> raw((size_t)n);
      for (int i = 0; i < n; i++) {
         raw[i] = reads[i].UP_Close;
         if (reads[i].hasCloseEnd()) CleanUniquePoints(reads[i].UP_Close);      // src/reader.cpp:272
      }
      searcher.SearchFarEnds(reads);            // src/pindel.cpp:1888
      std::vector<int32_t> o;
      for (int i = 0; i < n; i++) {
         const SPLIT_READ &r = reads[i];
         o.push_back(r.getUnmatchedSeq() == original[i] ? 0 : 1); o.push_back(r.getReadLength());
         o.push_back((int32_t)raw[i].size()); o.push_back((int32_t)r.UP_Close.size()); o.push_back((int32_t)r.UP_Far.size()); o.push_back(0);
         for (unsigned j = 0; j < raw[i].size(); j++) push(o, raw[i][j]);
         for (unsigned j = 0; j < r.UP_Close.size(); j++) push(o, r.UP_Close[j]);
         for (unsigned j = 0; j < r.UP_Far.size(); j++) push(o, r.UP_Far[j]);
      }
      std::ofstream g(argv[2], std::ios::binary);
      g.write((const char *)o.data(), (std::streamsize)(o.size() * sizeof(int32_t)));
      std::cerr << "host_mirror_test: " << n << " reads, " << searcher.kernelLaunches() << " kernel launches" << std::endl;
   } catch (const std::exception &e) {
      std::cerr << "host_mirror_test: " << e.what() << std::endl;
      return 1;
   }
   return 0;
}
