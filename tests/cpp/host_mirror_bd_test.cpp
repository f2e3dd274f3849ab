// tests/cpp/host_mirror_bd_test.cpp -- drives Searcher::loadRegion / getCorrespondingSearchWindowClusters / SearchFarEnds-without-
// clusters of include/pindel_pg.hpp the way main() drives BDData (src/pindel.cpp:1853, 1006), on inputs written by the pytest that
// launches it.  Links against libpindel_pg.so (GPU test) or the CPU emulation of the ABI (CPU test of the header's host logic).
//   host_mirror_bd_test <in.bin> <out.bin>
// in.bin : int32 n_chr, {int32 name_len, name bytes, uint32 padded length}*, int32 n_events, {int32 chr, uint32 pos, pos2, int32 chr,
//          uint32 pos, pos2}* (sorted), int32 window chr, uint32 start, end, int32 n_q, uint32 q*
// out.bin: int32 stream: per query [count, (chr, start, end)*]
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
   const int32_t n_chr = rd<int32_t>(f);
   std::vector<std::string> names, seqs;
   for (int c = 0; c < n_chr; c++) {
      const int32_t nl = rd<int32_t>(f);
      std::string name((size_t)nl, '\0');
      f.read(&name[0], nl);
      names.push_back(name);
      seqs.push_back(std::string((size_t)rd<uint32_t>(f), 'N'));      // only the lengths matter here
   }
   const int32_t n_ev = rd<int32_t>(f);
   std::vector<BreakDancerEvent> events((size_t)n_ev);
   for (int i = 0; i < n_ev; i++) {
      BreakDancerEvent &e = events[i];
      e.first.chromosome_id = rd<int32_t>(f); e.first.position = rd<uint32_t>(f); e.first.position2 = rd<uint32_t>(f);
      e.second.chromosome_id = rd<int32_t>(f); e.second.position = rd<uint32_t>(f); e.second.position2 = rd<uint32_t>(f);
   }
   SearchWindow w;
   w.chromosome_id = rd<int32_t>(f); w.start = rd<uint32_t>(f); w.end = rd<uint32_t>(f);
   const int32_t n_q = rd<int32_t>(f);
   std::vector<unsigned int> q((size_t)n_q);
   for (int i = 0; i < n_q; i++) q[i] = rd<uint32_t>(f);
   try {
      Searcher s;                                    // throws without a CUDA device when linked against the product
      s.setReference(seqs);
      s.loadRegion(events, names, w);                // g_bdData.loadRegion(currentWindow), src/pindel.cpp:1853
      const std::vector<SearchWindowCluster> cl = s.getCorrespondingSearchWindowClusters(q);
      s.unloadRegion();
      std::vector<int32_t> o;
      for (size_t i = 0; i < cl.size(); i++) {
         o.push_back((int32_t)cl[i].size());
         for (size_t k = 0; k < cl[i].size(); k++) { o.push_back(cl[i][k].chromosome_id); o.push_back((int32_t)cl[i][k].start); o.push_back((int32_t)cl[i][k].end); }
      }
      std::ofstream out(argv[2], std::ios::binary);
      out.write((const char *)o.data(), (std::streamsize)(o.size() * sizeof(int32_t)));
   } catch (const std::exception &e) {
      std::cerr << "host_mirror_bd_test: " << e.what() << std::endl;
      return 1;
   }
   return 0;
}
