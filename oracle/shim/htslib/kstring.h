/* oracle/shim/htslib/kstring.h -- TEST INFRASTRUCTURE ONLY (included, nothing used). */
#ifndef ORACLE_SHIM_HTSLIB_KSTRING_H
#define ORACLE_SHIM_HTSLIB_KSTRING_H
#include <stddef.h>
typedef struct { size_t l, m; char *s; } kstring_t;
#endif
