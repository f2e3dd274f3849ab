/* oracle/ref_hook_bd.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Hook around the reference's BreakDancer mask: BDData::loadRegion (src/bddata.cpp:850-946) and
 * BDData::getCorrespondingSearchWindowCluster (:949-967) on a caller-supplied event list.  The event list is a private
 * member that UpdateBD fills from the discordant pairs; to drive the two functions with arbitrary events (two positions
 * per coordinate, as UpdateBD builds them, src/bddata.cpp:706-707) this TU -- and only this TU -- sees BDData's members
 * as public.  Nothing of the algorithm is restated here: events in, the reference's own clusters out. */
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>
#define private public
#include "bddata.h"
#undef private
#include "pindel.h"
#include "control_state.h"

/* events: n rows of (first chromosome index, first pos, first pos2, second chromosome index, second pos, second pos2), any
 * order (sorted here with the reference's comparator, as UpdateBD does); window on chromosome `chr`; q: n_q values of
 * getLastAbsLocCloseEnd().  *out: per query [count, (chromosome index, start, end) x count].  Returns the stream length. */
extern "C" int64_t pgref_bd_clusters(int n_events, const int32_t *ev, int chr, uint32_t start, uint32_t end, int n_q, const uint32_t *q, int32_t **out)
{
   BDData bd;
   for (int i = 0; i < n_events; i++) {
      const int32_t *e = ev + 6 * i;
      BreakDancerCoordinate a(g_genome.getChr((unsigned)e[0])->getName(), (unsigned)e[1], (unsigned)e[2]);
      BreakDancerCoordinate b(g_genome.getChr((unsigned)e[3])->getName(), (unsigned)e[4], (unsigned)e[5]);
      bd.m_bdEvents.push_back(BreakDancerEvent(a, b));
   }
   std::sort(bd.m_bdEvents.begin(), bd.m_bdEvents.end(), sortOnFirstBDCoordinate);      /* src/bddata.cpp:809 */
   const Chromosome *c = g_genome.getChr((unsigned)chr);
   SearchWindow w(c, start, end);
   bd.loadRegion(w);
   std::vector<int32_t> s;
   for (int k = 0; k < n_q; k++) {
      SPLIT_READ r;
      r.UP_Close.push_back(UniquePoint(c, 20, q[k], FORWARD, SENSE, 0));
      const SearchWindowCluster &cl = bd.getCorrespondingSearchWindowCluster(r);
      s.push_back((int32_t)cl.size());
      for (unsigned j = 0; j < cl.size(); j++) {
         int id = -1;
         for (unsigned x = 0; g_genome.getChr(x) != NULL; x++) if (g_genome.getChr(x) == cl[j].getChromosome()) id = (int)x;
         s.push_back(id); s.push_back((int32_t)cl[j].getStart()); s.push_back((int32_t)cl[j].getEnd());
      }
   }
   *out = (int32_t *)malloc(sizeof(int32_t) * (s.size() ? s.size() : 1));
   if (!s.empty()) memcpy(*out, s.data(), sizeof(int32_t) * s.size());
   return (int64_t)s.size();
}
