/* oracle/shim/htslib/kseq.h -- TEST INFRASTRUCTURE ONLY (included, nothing used). */
#ifndef ORACLE_SHIM_HTSLIB_KSEQ_H
#define ORACLE_SHIM_HTSLIB_KSEQ_H
#endif
