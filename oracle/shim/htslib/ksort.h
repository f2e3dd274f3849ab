/* oracle/shim/htslib/ksort.h -- TEST INFRASTRUCTURE ONLY.  The reference instantiates
 * KSORT_INIT_GENERIC(uint32_t) (reader.cpp:34) but never calls the generated functions. */
#ifndef ORACLE_SHIM_HTSLIB_KSORT_H
#define ORACLE_SHIM_HTSLIB_KSORT_H
#define KSORT_INIT_GENERIC(type_t)
#endif
