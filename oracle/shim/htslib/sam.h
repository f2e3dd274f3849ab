/* oracle/shim/htslib/sam.h -- TEST INFRASTRUCTURE ONLY (not shipped, not linked into the product).
 *
 * The reference (genome/pindel) includes "htslib/sam.h" from pindel.h:34-35 and uses htslib's
 * BAM API in reader.cpp / bam2depth.cpp / search_MEI*.cpp.  htslib is not present in this
 * image, so this header declares the small subset of the public htslib API that those
 * translation units name.  The declarations follow htslib's documented public interface
 * (sam.h / hts.h); the definitions live in oracle/shim/minihts.cpp.
 */
#ifndef ORACLE_SHIM_HTSLIB_SAM_H
#define ORACLE_SHIM_HTSLIB_SAM_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
   int32_t n_targets;
   uint32_t *target_len;
   char **target_name;
   char *text;
   uint32_t l_text;
} bam_hdr_t;

typedef struct {
   int32_t tid;
   int32_t pos;
   uint16_t bin;
   uint8_t qual;
   uint8_t l_qname;
   uint16_t flag;
   uint16_t unused1;
   uint32_t n_cigar;
   int32_t l_qseq;
   int32_t mtid;
   int32_t mpos;
   int32_t isize;
} bam1_core_t;

typedef struct {
   bam1_core_t core;
   int l_data, m_data;
   uint8_t *data;
} bam1_t;

typedef struct samFile_s samFile;
typedef samFile htsFile;
typedef struct hts_idx_s hts_idx_t;
typedef struct hts_itr_s hts_itr_t;

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CBACK       9
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
#define BAM_CIGAR_TYPE  0x3C1A7
#define bam_cigar_op(c) ((c)&BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c)>>BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (BAM_CIGAR_TYPE>>((o)<<1)&3)

#define bam_is_rev(b) (((b)->core.flag&BAM_FREVERSE) != 0)
#define bam_is_mrev(b) (((b)->core.flag&BAM_FMREVERSE) != 0)
#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (((b)->core.l_qseq + 1)>>1))
#define bam_get_aux(b)   ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (((b)->core.l_qseq + 1)>>1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar<<2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1)>>1))
#define bam_seqi(s, i) ((s)[(i)>>1] >> ((~(i)&1)<<2) & 0xf)

extern const char seq_nt16_str[];

samFile *sam_open(const char *fn, const char *mode);
int sam_close(samFile *fp);
bam_hdr_t *sam_hdr_read(samFile *fp);
void bam_hdr_destroy(bam_hdr_t *h);
int bam_name2id(bam_hdr_t *h, const char *ref);
hts_idx_t *sam_index_load(samFile *fp, const char *fn);
void hts_idx_destroy(hts_idx_t *idx);
hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, int beg, int end);
hts_itr_t *sam_itr_querys(const hts_idx_t *idx, bam_hdr_t *hdr, const char *region);
int sam_itr_next(samFile *fp, hts_itr_t *itr, bam1_t *b);
void hts_itr_destroy(hts_itr_t *itr);
int sam_read1(samFile *fp, bam_hdr_t *h, bam1_t *b);
#define bam_itr_destroy(i) hts_itr_destroy(i)

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
bam1_t *bam_dup1(const bam1_t *b);
bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src);
int32_t bam_endpos(const bam1_t *b);

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int32_t bam_aux2i(const uint8_t *s);
char bam_aux2A(const uint8_t *s);
char *bam_aux2Z(const uint8_t *s);

/* pileup API named by bam2depth.cpp (dead code in the reference: pindel.cpp:1734,1740,2007) */
typedef struct {
   bam1_t *b;
   int32_t qpos;
   int indel, level;
   uint32_t is_del:1, is_head:1, is_tail:1, is_refskip:1, aux:28;
} bam_pileup1_t;
typedef int (*bam_plp_auto_f)(void *data, bam1_t *b);
typedef struct bam_mplp_s *bam_mplp_t;
bam_mplp_t bam_mplp_init(int n, bam_plp_auto_f func, void **data);
void bam_mplp_destroy(bam_mplp_t iter);
void bam_mplp_set_maxcnt(bam_mplp_t iter, int maxcnt);
int bam_mplp_auto(bam_mplp_t iter, int *_tid, int *_pos, int *n_plp, const bam_pileup1_t **plp);

#ifdef __cplusplus
}
#endif
#endif
