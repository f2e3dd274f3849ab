/* oracle/shim/hts_stub.cpp -- TEST INFRASTRUCTURE ONLY.
 * Aborting definitions for the htslib entry points declared in shim/htslib/sam.h.  The
 * oracle build drives the reference with Pindel-text input (-p/-P), which never reaches
 * them (reader.cpp:1428-1486 dispatches on the input kind); BAM input is out of reach
 * until a BGZF/BAM reader exists.
 */
#include <cstdio>
#include <cstdlib>
#include "htslib/sam.h"

static void shim_die(const char *what)
{
   std::fprintf(stderr, "oracle shim: %s called, but htslib is not available in this build (text input only)\n", what);
   std::abort();
}

extern "C" {
const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
samFile *sam_open(const char *, const char *) { shim_die("sam_open"); return 0; }
int sam_close(samFile *) { shim_die("sam_close"); return -1; }
bam_hdr_t *sam_hdr_read(samFile *) { shim_die("sam_hdr_read"); return 0; }
void bam_hdr_destroy(bam_hdr_t *) { shim_die("bam_hdr_destroy"); }
int bam_name2id(bam_hdr_t *, const char *) { shim_die("bam_name2id"); return -1; }
hts_idx_t *sam_index_load(samFile *, const char *) { shim_die("sam_index_load"); return 0; }
void hts_idx_destroy(hts_idx_t *) { shim_die("hts_idx_destroy"); }
hts_itr_t *sam_itr_queryi(const hts_idx_t *, int, int, int) { shim_die("sam_itr_queryi"); return 0; }
hts_itr_t *sam_itr_querys(const hts_idx_t *, bam_hdr_t *, const char *) { shim_die("sam_itr_querys"); return 0; }
int sam_itr_next(samFile *, hts_itr_t *, bam1_t *) { shim_die("sam_itr_next"); return -1; }
void hts_itr_destroy(hts_itr_t *) { shim_die("hts_itr_destroy"); }
int sam_read1(samFile *, bam_hdr_t *, bam1_t *) { shim_die("sam_read1"); return -1; }
bam1_t *bam_init1(void) { shim_die("bam_init1"); return 0; }
void bam_destroy1(bam1_t *) { shim_die("bam_destroy1"); }
bam1_t *bam_dup1(const bam1_t *) { shim_die("bam_dup1"); return 0; }
bam1_t *bam_copy1(bam1_t *, const bam1_t *) { shim_die("bam_copy1"); return 0; }
int32_t bam_endpos(const bam1_t *) { shim_die("bam_endpos"); return 0; }
uint8_t *bam_aux_get(const bam1_t *, const char[2]) { shim_die("bam_aux_get"); return 0; }
int32_t bam_aux2i(const uint8_t *) { shim_die("bam_aux2i"); return 0; }
char bam_aux2A(const uint8_t *) { shim_die("bam_aux2A"); return 0; }
char *bam_aux2Z(const uint8_t *) { shim_die("bam_aux2Z"); return 0; }
bam_mplp_t bam_mplp_init(int, bam_plp_auto_f, void **) { shim_die("bam_mplp_init"); return 0; }
void bam_mplp_destroy(bam_mplp_t) { shim_die("bam_mplp_destroy"); }
void bam_mplp_set_maxcnt(bam_mplp_t, int) { shim_die("bam_mplp_set_maxcnt"); }
int bam_mplp_auto(bam_mplp_t, int *, int *, int *, const bam_pileup1_t **) { shim_die("bam_mplp_auto"); return -1; }
}
