/* oracle/shim/absfix.h -- TEST INFRASTRUCTURE ONLY.  Force-included when compiling the
 * reference: g++ >= 7 rejects abs(unsigned - unsigned) as ambiguous (pindel.cpp:1557,
 * bddata.cpp:122, genotyping.cpp:127); older compilers resolved it to abs(int). */
#ifndef ORACLE_SHIM_ABSFIX_H
#define ORACLE_SHIM_ABSFIX_H
#ifdef __cplusplus
#include <cstdlib>
#include <cmath>
#include <stdlib.h>
#include <math.h>
inline int abs(unsigned int x) { return ::abs(static_cast<int>(x)); }
#endif
#endif
