/* oracle/shim/minihts.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A small, from-scratch reader for coordinate-sorted BAM files over zlib, offering the slice of the public htslib API
 * that the reference names in reader.cpp (sam_open, sam_hdr_read, bam_name2id, sam_index_load, sam_itr_queryi,
 * sam_itr_next, bam_init1/dup1/destroy1, bam_aux_get/2i/2A/2Z ...).  htslib itself is not installed in this image; this
 * lets the oracle build of the reference (oracle/_ref) run its BAM scenarios -- demo/RUNME's simulated_config.txt and
 * devtools/run_regressiontests.sh:7 -- so that they can serve as parity cases.
 *
 * Written from the SAM/BAM format specification (BGZF = concatenated gzip members; BAM = "BAM\1", header text,
 * reference dictionary, then records of a 32-byte fixed part + qname/cigar/seq/qual/aux), not from htslib's sources.
 * The .bai index is not parsed: a query walks the in-memory record table of the whole (small) file and applies the
 * region test the indexed iterator applies (same chromosome, pos < end, endpos > beg), in file order.
 * The pileup API (bam2depth.cpp, dead code in the reference: src/pindel.cpp:1734,1740,2007) still aborts.
 */
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "htslib/sam.h"

namespace {
void die(const char *what)
{
   std::fprintf(stderr, "oracle minihts: %s\n", what);
   std::abort();
}

inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

struct Rec {
   size_t off;       /* offset of block_size field in the inflated stream */
   int32_t tid, pos, end;
};
} // namespace

struct samFile_s {
   std::vector<uint8_t> data;    /* whole file, inflated */
   size_t first_record;
   std::vector<Rec> recs;
   bool indexed;
};
struct hts_idx_s { int dummy; };
struct hts_itr_s {
   int tid, beg, end;
   size_t next;      /* index into recs */
   bool done;
};

static int32_t ref_span(const uint8_t *rec)
{
   /* rec points at refID (after block_size) */
   uint8_t l_qname = rec[8];
   uint16_t n_cigar = le16(rec + 12);
   uint16_t flag = le16(rec + 14);
   const uint8_t *cig = rec + 32 + l_qname;
   int32_t span = 0;
   if (!(flag & BAM_FUNMAP)) {
      for (unsigned k = 0; k < n_cigar; k++) {
         uint32_t c = le32(cig + 4 * k);
         uint32_t op = c & BAM_CIGAR_MASK, len = c >> BAM_CIGAR_SHIFT;
         if (op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF) span += (int32_t)len;
      }
   }
   return span > 0 ? span : 1;       /* unmapped or no reference-consuming operation: occupies one base */
}

extern "C" {

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

/* optional replacement of the record walk in sam_hdr_read: the table of the stream the inflate hook produced last
   (24-byte rows: offset u64, refID i32, pos i32, end i32, 4 bytes the reader does not use); 0 = success */
int (*minihts_index_hook)(size_t first_record, std::vector<uint8_t> &rows24) = NULL;

/* optional replacement of the zlib loop in sam_open: whole BGZF file in, inflated stream out; 0 = success */
int (*minihts_inflate_hook)(const uint8_t *file, size_t len, std::vector<uint8_t> &out) = NULL;

samFile *sam_open(const char *fn, const char *)
{
   FILE *f = std::fopen(fn, "rb");
   if (!f) return NULL;
   std::vector<uint8_t> raw;
   uint8_t buf[1 << 16];
   size_t n;
   while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) raw.insert(raw.end(), buf, buf + n);
   std::fclose(f);
   samFile *fp = new samFile_s();
   fp->indexed = false;
   if (minihts_inflate_hook) {
      /* the drop-in demonstration installs the device inflater here (oracle/ref_gpu_bridge.cpp -> pg_bgzf_inflate);
         the pure reference build never sets the hook and inflates with zlib below */
      raw.resize(raw.size() + 16);          /* readable padding, not part of the file */
      if (minihts_inflate_hook(raw.data(), raw.size() - 16, fp->data) != 0) die("inflate hook failed");
      if (fp->data.size() < 12 || std::memcmp(fp->data.data(), "BAM\1", 4) != 0) die("not a BAM file");
      return fp;
   }
   /* BGZF: a series of gzip members */
   z_stream zs;
   std::memset(&zs, 0, sizeof zs);
   if (inflateInit2(&zs, 15 + 32) != Z_OK) die("inflateInit2");
   zs.next_in = raw.data();
   zs.avail_in = (uInt)raw.size();
   std::vector<uint8_t> out(1 << 20);
   while (zs.avail_in > 0) {
      zs.next_out = out.data();
      zs.avail_out = (uInt)out.size();
      int rc = inflate(&zs, Z_NO_FLUSH);
      if (rc != Z_OK && rc != Z_STREAM_END) die("inflate failed (not a BGZF/BAM file?)");
      fp->data.insert(fp->data.end(), out.data(), out.data() + (out.size() - zs.avail_out));
      if (rc == Z_STREAM_END) {
         if (zs.avail_in == 0) break;
         if (inflateReset(&zs) != Z_OK) die("inflateReset");
      }
   }
   inflateEnd(&zs);
   if (fp->data.size() < 12 || std::memcmp(fp->data.data(), "BAM\1", 4) != 0) die("not a BAM file");
   return fp;
}

int sam_close(samFile *fp) { delete fp; return 0; }

bam_hdr_t *sam_hdr_read(samFile *fp)
{
   const uint8_t *d = fp->data.data();
   size_t p = 4;
   bam_hdr_t *h = (bam_hdr_t *)std::calloc(1, sizeof(bam_hdr_t));
   h->l_text = le32(d + p); p += 4;
   h->text = (char *)std::malloc(h->l_text + 1);
   std::memcpy(h->text, d + p, h->l_text); h->text[h->l_text] = 0; p += h->l_text;
   h->n_targets = (int32_t)le32(d + p); p += 4;
   h->target_name = (char **)std::calloc((size_t)h->n_targets, sizeof(char *));
   h->target_len = (uint32_t *)std::calloc((size_t)h->n_targets, sizeof(uint32_t));
   for (int i = 0; i < h->n_targets; i++) {
      uint32_t l = le32(d + p); p += 4;
      h->target_name[i] = (char *)std::malloc(l);
      std::memcpy(h->target_name[i], d + p, l); p += l;
      h->target_len[i] = le32(d + p); p += 4;
   }
   fp->first_record = p;
   if (!fp->indexed && minihts_index_hook && minihts_inflate_hook) {
      std::vector<uint8_t> rows;
      if (minihts_index_hook(p, rows) != 0) die("index hook failed");
      fp->recs.resize(rows.size() / 24);
      for (size_t i = 0; i < fp->recs.size(); i++) {
         const uint8_t *r = rows.data() + 24 * i;
         uint64_t off;
         std::memcpy(&off, r, 8);
         fp->recs[i].off = (size_t)off;
         fp->recs[i].tid = (int32_t)le32(r + 8); fp->recs[i].pos = (int32_t)le32(r + 12); fp->recs[i].end = (int32_t)le32(r + 16);
      }
      fp->indexed = true;
   }
   if (!fp->indexed) {
      /* record table for region queries */
      size_t q = p, n = fp->data.size();
      while (q + 4 <= n) {
         uint32_t bs = le32(d + q);
         if (q + 4 + bs > n) die("truncated BAM record");
         const uint8_t *rec = d + q + 4;
         Rec r;
         r.off = q; r.tid = (int32_t)le32(rec); r.pos = (int32_t)le32(rec + 4);
         r.end = r.pos + ref_span(rec);
         fp->recs.push_back(r);
         q += 4 + bs;
      }
      fp->indexed = true;
   }
   return h;
}

void bam_hdr_destroy(bam_hdr_t *h)
{
   if (!h) return;
   for (int i = 0; i < h->n_targets; i++) std::free(h->target_name[i]);
   std::free(h->target_name); std::free(h->target_len); std::free(h->text); std::free(h);
}

int bam_name2id(bam_hdr_t *h, const char *ref)
{
   for (int i = 0; i < h->n_targets; i++) if (std::strcmp(h->target_name[i], ref) == 0) return i;
   return -1;
}

hts_idx_t *sam_index_load(samFile *, const char *) { return new hts_idx_s(); }
void hts_idx_destroy(hts_idx_t *idx) { delete idx; }

hts_itr_t *sam_itr_queryi(const hts_idx_t *, int tid, int beg, int end)
{
   hts_itr_t *it = new hts_itr_s();
   it->tid = tid; it->beg = beg < 0 ? 0 : beg; it->end = end; it->next = 0; it->done = false;
   return it;
}
hts_itr_t *sam_itr_querys(const hts_idx_t *, bam_hdr_t *, const char *) { die("sam_itr_querys is not implemented"); return NULL; }
void hts_itr_destroy(hts_itr_t *it) { delete it; }

static void fill(bam1_t *b, const uint8_t *p /* at block_size */)
{
   uint32_t bs = le32(p);
   const uint8_t *rec = p + 4;
   b->core.tid = (int32_t)le32(rec);
   b->core.pos = (int32_t)le32(rec + 4);
   b->core.l_qname = rec[8];
   b->core.qual = rec[9];
   b->core.bin = le16(rec + 10);
   b->core.n_cigar = le16(rec + 12);
   b->core.flag = le16(rec + 14);
   b->core.l_qseq = (int32_t)le32(rec + 16);
   b->core.mtid = (int32_t)le32(rec + 20);
   b->core.mpos = (int32_t)le32(rec + 24);
   b->core.isize = (int32_t)le32(rec + 28);
   int l = (int)bs - 32;
   if (l > b->m_data) {
      b->m_data = l + 64;
      b->data = (uint8_t *)std::realloc(b->data, (size_t)b->m_data);
   }
   b->l_data = l;
   std::memcpy(b->data, rec + 32, (size_t)l);
}

int sam_itr_next(samFile *fp, hts_itr_t *it, bam1_t *b)
{
   if (it->done) return -1;
   if (!fp->indexed) die("sam_itr_next before sam_hdr_read");
   while (it->next < fp->recs.size()) {
      const Rec &r = fp->recs[it->next++];
      if (r.tid != it->tid) {
         if (r.tid > it->tid || r.tid < 0) { it->done = true; return -1; }   /* coordinate-sorted: past the chromosome */
         continue;
      }
      if (r.pos >= it->end) { it->done = true; return -1; }
      if (r.end > it->beg) {
         fill(b, fp->data.data() + r.off);
         return 0;
      }
   }
   it->done = true;
   return -1;
}

int sam_read1(samFile *, bam_hdr_t *, bam1_t *) { die("sam_read1 is not implemented (only bam2depth.cpp, dead code, uses it)"); return -1; }

bam1_t *bam_init1(void) { return (bam1_t *)std::calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (b) { std::free(b->data); std::free(b); } }
bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src)
{
   uint8_t *data = dst->data;
   int m = dst->m_data;
   if (m < src->l_data) { m = src->l_data + 64; data = (uint8_t *)std::realloc(data, (size_t)m); }
   std::memcpy(data, src->data, (size_t)src->l_data);
   *dst = *src;
   dst->m_data = m; dst->data = data;
   return dst;
}
bam1_t *bam_dup1(const bam1_t *src)
{
   if (!src) return NULL;
   bam1_t *b = bam_init1();
   return bam_copy1(b, src);
}
int32_t bam_endpos(const bam1_t *b)
{
   int32_t span = 0;
   if (!(b->core.flag & BAM_FUNMAP)) {
      const uint32_t *cig = bam_get_cigar(b);
      for (unsigned k = 0; k < b->core.n_cigar; k++) {
         uint32_t op = cig[k] & BAM_CIGAR_MASK;
         if (op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF) span += (int32_t)(cig[k] >> BAM_CIGAR_SHIFT);
      }
   }
   return b->core.pos + (span > 0 ? span : 1);
}

static int aux_size(uint8_t type, const uint8_t *v, const uint8_t *end)
{
   switch (type) {
   case 'A': case 'c': case 'C': return 1;
   case 's': case 'S': return 2;
   case 'i': case 'I': case 'f': return 4;
   case 'd': return 8;
   case 'Z': case 'H': { const uint8_t *q = v; while (q < end && *q) q++; return (int)(q - v) + 1; }
   case 'B': {
      if (v + 5 > end) return -1;
      int es = aux_size(v[0], NULL, NULL);
      return 5 + es * (int)le32(v + 1);
   }
   default: return -1;
   }
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2])
{
   uint8_t *s = bam_get_aux(b), *end = b->data + b->l_data;
   while (s + 3 <= end) {
      uint8_t *type = s + 2;
      if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return type;
      int sz = aux_size(*type, type + 1, end);
      if (sz < 0) return NULL;
      s = type + 1 + sz;
   }
   return NULL;
}
int32_t bam_aux2i(const uint8_t *s)
{
   uint8_t t = *s++;
   switch (t) {
   case 'c': return (int8_t)s[0];
   case 'C': return s[0];
   case 's': return (int16_t)le16(s);
   case 'S': return le16(s);
   case 'i': return (int32_t)le32(s);
   case 'I': return (int32_t)le32(s);
   default: return 0;
   }
}
char bam_aux2A(const uint8_t *s) { return (*s == 'A') ? (char)s[1] : 0; }
char *bam_aux2Z(const uint8_t *s) { return (*s == 'Z' || *s == 'H') ? (char *)(s + 1) : NULL; }

bam_mplp_t bam_mplp_init(int, bam_plp_auto_f, void **) { die("pileup API is not implemented (dead code in the reference)"); return 0; }
void bam_mplp_destroy(bam_mplp_t) { die("pileup API is not implemented"); }
void bam_mplp_set_maxcnt(bam_mplp_t, int) { die("pileup API is not implemented"); }
int bam_mplp_auto(bam_mplp_t, int *, int *, int *, const bam_pileup1_t **) { die("pileup API is not implemented"); return -1; }

} // extern "C"
