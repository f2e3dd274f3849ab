// tests/emul/abi_emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The entry points of include/pindel_pg.h that the drop-in bridge (oracle/ref_gpu_bridge.cpp) binds, served by the CPU
// emulators of the device programs (pg_emul.cpp: the warp program of pg_warp.cuh on 32 fibers; sv_emul.cpp: pg_sv.cuh as
// plain C++).  It exists so the *host logic* of the bridge -- marshalling of SPLIT_READs, replay of the classification
// records, the reporter calls -- can be exercised in the GPU-less container (`pytest -m "not gpu"`), by linking the very
// same bridge object against this library instead of libpindel_pg.so (oracle/_ref/pindel_emul_demo).  It is not a
// fallback of the product: libpindel_pg.so contains no host implementation and nothing in pindel_b200/ loads this file.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "pg_emul.cpp"
#include "sv_emul.cpp"
#include "rp_emul.cpp"
#include "bgzf_emul.cpp"
#include "bam_emul.cpp"
#include "boxsort_emul.cpp"
#include "sr_emul.cpp"
#include "bd_emul.cpp"
#include "../../pindel_b200/csrc/pg_host_helpers.h"

struct pg_ctx {
   pg_settings st;
   std::string err;
   std::vector<std::string> chr;
   std::vector<uint8_t> params;
   uint32_t table[PG_TABLE];
   std::vector<pg_read_out> out;
   std::vector<pg_run> runs;
   uint64_t launches = 0, sv_launches = 0, rp_launches = 0, bx_launches = 0;
   // copy of the last far-end-only batch ("resident")
   int res_n = -1;
   std::vector<uint8_t> res_seq, res_md;
   std::vector<uint32_t> res_off;
   std::vector<int16_t> res_isz;
   std::vector<int32_t> res_chr;
   PgBdPrepared bd;               // the loaded BreakDancer region (pg_bd_load_region)
   bool bd_loaded = false;
   uint64_t bd_lookups = 0;
   std::vector<uint8_t> stream;   // the last inflated stream
   std::vector<PgBamRec> bam_table;   // its record table (pg_bam_index with a buffer)
};

static std::string g_create_error;

static int fail(pg_ctx *c, int code, const char *msg)
{
   if (c) c->err = msg;
   return code;
}

extern "C" {

void pg_default_settings(pg_settings *s)
{
   memset(s, 0, sizeof *s);
   s->seq_error_rate = 0.01;
   s->sensitivity = 0.95;
   s->max_mismatch_rate = 0.02;
   s->additional_mismatch = 1;
   s->min_close = 8;
   s->min_perfect_match = 3;
   s->max_range_index = 2;
}

const char *pg_create_error(void) { return g_create_error.c_str(); }
const char *pg_last_error(pg_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

pg_ctx *pg_create(const pg_settings *s)
{
   pg_ctx *c = new pg_ctx;
   c->st = *s;
   c->params.resize(sizeof(PgParams));
   if (pgemu_make_params(c->params.data(), s->seq_error_rate, s->sensitivity, s->max_mismatch_rate, s->additional_mismatch,
                         s->min_close, s->min_perfect_match, s->max_range_index) != 0) {
      g_create_error = "bad settings";
      delete c;
      return NULL;
   }
   pg_prob_table(0.001 + s->seq_error_rate, s->sensitivity, c->table);
   return c;
}

void pg_destroy(pg_ctx *ctx) { delete ctx; }

int pg_max_mismatch_table(pg_ctx *ctx, uint32_t *out)
{
   memcpy(out, ctx->table, sizeof(uint32_t) * PG_TABLE);
   return PG_OK;
}

int pg_set_reference(pg_ctx *ctx, int32_t n_chr, const char *const *seq, const uint32_t *len)
{
   ctx->chr.clear();
   for (int i = 0; i < n_chr; i++) ctx->chr.push_back(std::string(seq[i], len[i]));
   return PG_OK;
}

static int run_inner(pg_ctx *ctx, const pg_reads *in, pg_results *res, int mode);
static int run(pg_ctx *ctx, const pg_reads *in, pg_results *res, int mode)
{
   /* a far-end batch without regions takes every read's cluster from the loaded BreakDancer region */
   if (mode == 2 && ctx->bd_loaded && !(in->regions && in->region_off) && in->close_info && in->n > 0) {
      PgBdRegionView R;
      R.start = ctx->bd.start.data(); R.drop = ctx->bd.drop.data(); R.sel = ctx->bd.sel.data(); R.m = (uint32_t)ctx->bd.start.size();
      R.z = ctx->bd.z; R.ws = ctx->bd.ws; R.we = ctx->bd.we;
      std::vector<uint32_t> off((size_t)in->n + 1, 0);
      std::vector<pg_region> reg;
      for (int i = 0; i < in->n; i++) {
         off[i] = (uint32_t)reg.size();
         ctx->bd_lookups++;
         if (in->close_info[i].max_len == 0) continue;
         uint32_t S, E;
         pgbd_range(R, in->close_info[i].last_abs_loc, S, E);
         if (E > S) {
            const uint32_t c = pgbd_cluster(R, S, E, NULL, 0);
            reg.resize(reg.size() + c);
            pgbd_cluster(R, S, E, reinterpret_cast<PgRegion *>(&reg[off[i]]), c);
         }
      }
      off[in->n] = (uint32_t)reg.size();
      if (reg.empty()) reg.resize(1);
      pg_reads in2 = *in;
      in2.regions = reg.data(); in2.region_off = off.data();
      return run_inner(ctx, &in2, res, mode);
   }
   return run_inner(ctx, in, res, mode);
}
static int run_inner(pg_ctx *ctx, const pg_reads *in, pg_results *res, int mode)
{
   if (ctx->chr.empty()) return fail(ctx, PG_ERR_STATE, "no reference");
   const int n = in->n;
   std::vector<const char *> ptr;
   std::vector<uint32_t> len;
   for (size_t i = 0; i < ctx->chr.size(); i++) {
      ptr.push_back(ctx->chr[i].data());
      len.push_back((uint32_t)ctx->chr[i].size());
   }
   ctx->out.assign((size_t)n, pg_read_out());
   for (uint32_t cap = (uint32_t)n * 16 + 1024;; cap *= 4) {
      ctx->runs.assign(cap, pg_run());
      uint32_t used = 0;
      const int of = pgemu_search((int)ptr.size(), ptr.data(), len.data(), ctx->params.data(), n, mode, (const uint8_t *)in->seq_blob,
                                  in->seq_off, (const uint8_t *)in->matched_d, in->matched_rel_pos, in->insert_size, in->chr_id,
                                  in->close_info, in->regions, in->region_off, ctx->out.data(), ctx->runs.data(), cap, &used);
      if (!of) {
         ctx->runs.resize(used);
         break;
      }
   }
   ctx->launches++;
   ctx->res_n = -1;
   if (mode == 2 && in->matched_d && in->insert_size) {
      ctx->res_n = n;
      ctx->res_seq.assign((const uint8_t *)in->seq_blob, (const uint8_t *)in->seq_blob + in->seq_off[n]);
      ctx->res_off.assign(in->seq_off, in->seq_off + n + 1);
      ctx->res_md.assign((const uint8_t *)in->matched_d, (const uint8_t *)in->matched_d + n);
      ctx->res_isz.assign(in->insert_size, in->insert_size + n);
      ctx->res_chr.assign((size_t)n, 0);
      if (in->chr_id) ctx->res_chr.assign(in->chr_id, in->chr_id + n);
   }
   return pg_download(ctx, res);
}

int pg_download(pg_ctx *ctx, pg_results *res)
{
   res->n_runs = (uint32_t)ctx->runs.size();
   if (res->n_runs > res->runs_cap) return fail(ctx, PG_ERR_RUNS_CAP, "run buffer too small");
   if (!ctx->out.empty()) memcpy(res->out, ctx->out.data(), sizeof(pg_read_out) * ctx->out.size());
   if (!ctx->runs.empty()) memcpy(res->runs, ctx->runs.data(), sizeof(pg_run) * ctx->runs.size());
   return PG_OK;
}

int pg_close_end(pg_ctx *ctx, const pg_reads *in, pg_results *res) { return run(ctx, in, res, 1); }
int pg_far_end(pg_ctx *ctx, const pg_reads *in, pg_results *res) { return run(ctx, in, res, 2); }
int pg_search(pg_ctx *ctx, const pg_reads *in, pg_results *res) { return run(ctx, in, res, 3); }
uint64_t pg_kernel_launches(pg_ctx *ctx) { return ctx ? ctx->launches : 0; }

/* pg_group on the emulators: the same dealing of contiguous ranges and the same merge (run offsets made global), serially */
struct pg_group {
   std::vector<pg_ctx *> ctx;
   std::string err;
};
pg_group *pg_group_create(const pg_settings *s, int32_t n_devices, const int32_t *)
{
   pg_group *g = new pg_group;
   for (int k = 0; k < n_devices; k++) g->ctx.push_back(pg_create(s));
   return g;
}
void pg_group_destroy(pg_group *g) { if (g) { for (pg_ctx *c : g->ctx) pg_destroy(c); delete g; } }
const char *pg_group_last_error(pg_group *g) { return g ? g->err.c_str() : "null group"; }
int32_t pg_group_size(pg_group *g) { return g ? (int32_t)g->ctx.size() : 0; }
pg_ctx *pg_group_context(pg_group *g, int32_t k) { return g->ctx[k]; }
int pg_group_set_reference(pg_group *g, int32_t n_chr, const char *const *seq, const uint32_t *len)
{
   for (pg_ctx *c : g->ctx) pg_set_reference(c, n_chr, seq, len);
   return PG_OK;
}
static int group_run(pg_group *g, const pg_reads *in, pg_results *res, int mode)
{
   const int nd = (int)g->ctx.size(), n = in->n;
   res->n_runs = 0;
   std::vector<std::vector<pg_run>> runs(nd);
   std::vector<uint32_t> nruns(nd, 0);
   for (int k = 0; k < nd; k++) {
      const int a = (int)((long long)n * k / nd), b = (int)((long long)n * (k + 1) / nd), m = b - a;
      if (m == 0) continue;
      pg_reads r = *in;
      r.n = m;
      std::vector<uint32_t> off(m + 1), roff(m + 1);
      for (int i = 0; i <= m; i++) off[i] = in->seq_off[a + i] - in->seq_off[a];
      r.seq_blob = in->seq_blob + in->seq_off[a]; r.seq_off = off.data();
      if (in->matched_d) r.matched_d = in->matched_d + a;
      if (in->matched_rel_pos) r.matched_rel_pos = in->matched_rel_pos + a;
      if (in->insert_size) r.insert_size = in->insert_size + a;
      if (in->chr_id) r.chr_id = in->chr_id + a;
      if (in->close_info) r.close_info = in->close_info + a;
      if (in->regions && in->region_off) {
         for (int i = 0; i <= m; i++) roff[i] = in->region_off[a + i] - in->region_off[a];
         r.regions = in->regions + in->region_off[a]; r.region_off = roff.data();
      }
      runs[k].resize((size_t)m * 16 + 1024);
      pg_results pr;
      pr.out = res->out + a; pr.runs = runs[k].data(); pr.runs_cap = (uint32_t)runs[k].size(); pr.n_runs = 0;
      int rc = run(g->ctx[k], &r, &pr, mode);
      if (rc == PG_ERR_RUNS_CAP) { runs[k].resize(pr.n_runs); pr.runs = runs[k].data(); pr.runs_cap = pr.n_runs; rc = run(g->ctx[k], &r, &pr, mode); }
      if (rc != PG_OK) { g->err = g->ctx[k]->err; return rc; }
      nruns[k] = pr.n_runs;
      for (int i = 0; i < m; i++) res->out[a + i].run_off += res->n_runs;
      res->n_runs += pr.n_runs;
   }
   if (res->n_runs > res->runs_cap) { g->err = "run buffer too small"; return PG_ERR_RUNS_CAP; }
   uint32_t p = 0;
   for (int k = 0; k < nd; k++) { if (nruns[k]) memcpy(res->runs + p, runs[k].data(), sizeof(pg_run) * nruns[k]); p += nruns[k]; }
   return PG_OK;
}
int pg_group_close_end(pg_group *g, const pg_reads *in, pg_results *res) { return group_run(g, in, res, 1); }
int pg_group_far_end(pg_group *g, const pg_reads *in, pg_results *res) { return group_run(g, in, res, 2); }
int pg_group_search(pg_group *g, const pg_reads *in, pg_results *res) { return group_run(g, in, res, 3); }

int pg_classify(pg_ctx *ctx, const pg_sv_reads *in, const pg_sv_window *win, pg_sv_results *res)
{
   res->n_records = 0;
   if (in->n_reads == 0) return PG_OK;
   const uint32_t w[11] = {win->driver_mask, 0, win->window_end, win->num_boxes, win->box_size, win->region_start_defined,
                           win->region_start, win->region_end_defined, win->region_end, (uint32_t)win->min_inversion_size,
                           (uint32_t)win->min_matched_bases};
   const std::string &c = ctx->chr[(size_t)win->chr];
   std::vector<PgSvRecord> rec((size_t)in->n_reads * 4 + 1024);
   const int64_t k = svemu_classify(c.data(), (uint32_t)c.size(), ctx->st.seq_error_rate, ctx->st.sensitivity, w, in->n_reads, in->seq_blob,
                                    in->seq_off, in->matched_d, in->insert_size, in->rflags, in->run_off, in->n_close_runs,
                                    (const PgRun *)in->runs, rec.data(), (uint32_t)rec.size());
   ctx->sv_launches++;
   res->n_records = (uint32_t)k;
   if (res->n_records > res->records_cap) return fail(ctx, PG_ERR_RUNS_CAP, "record buffer too small");
   memcpy(res->records, rec.data(), sizeof(pg_sv_record) * (size_t)k);
   std::sort(res->records, res->records + k, [](const pg_sv_record &a, const pg_sv_record &b) {
      if (a.driver != b.driver) return a.driver < b.driver;
      if (a.read != b.read) return a.read < b.read;
      return a.seq < b.seq;
   });
   return PG_OK;
}

// after a far-end-only call: the batch kept above + the caller's close runs, re-assembled into the host-array form
int pg_classify_resident(pg_ctx *ctx, const pg_sv_window *win, const uint8_t *rflags, const uint32_t *close_run_off,
                         const pg_run *close_runs, pg_sv_results *res)
{
   if (ctx->res_n < 0) return fail(ctx, PG_ERR_STATE, "no far-end launch is resident");
   const int n = ctx->res_n;
   std::vector<uint32_t> run_off((size_t)n + 1, 0);
   std::vector<uint16_t> n_close((size_t)n);
   std::vector<uint8_t> flags((size_t)n);
   std::vector<pg_run> runs;
   for (int i = 0; i < n; i++) {
      const uint32_t c0 = close_run_off[i], c1 = close_run_off[i + 1];
      runs.insert(runs.end(), close_runs + c0, close_runs + c1);
      n_close[i] = (uint16_t)(c1 - c0);
      const pg_read_out &o = ctx->out[i];
      const pg_run *fr = ctx->runs.data() + o.run_off + o.n_close_runs;
      runs.insert(runs.end(), fr, fr + o.n_far_runs);
      run_off[i + 1] = (uint32_t)runs.size();
      flags[i] = (uint8_t)((rflags ? rflags[i] & PG_SVR_USED : 0) | ((o.n_far_runs && fr[0].chr != ctx->res_chr[i]) ? PG_SVR_OTHERCHR : 0));
   }
   pg_sv_reads in;
   memset(&in, 0, sizeof in);
   in.n_reads = n;
   in.seq_blob = ctx->res_seq.data();
   in.seq_off = ctx->res_off.data();
   in.matched_d = ctx->res_md.data();
   in.insert_size = ctx->res_isz.data();
   in.rflags = flags.data();
   in.run_off = run_off.data();
   in.n_close_runs = n_close.data();
   in.runs = runs.data();
   return pg_classify(ctx, &in, win, res);
}

int pg_rp_modify(pg_ctx *ctx, int32_t n, const pg_rp *in, pg_rp_cluster *out)
{
   rpemu_modify(n, (const PgRp *)in, (PgRpCluster *)out);
   ctx->rp_launches++;
   return PG_OK;
}
int pg_rp_summarize(pg_ctx *ctx, int32_t n, const pg_rp *in, pg_rp_cluster *inout)
{
   rpemu_summarize(n, (const PgRp *)in, (PgRpCluster *)inout);
   ctx->rp_launches++;
   return PG_OK;
}
int pg_rp_modify_interchr(pg_ctx *ctx, int32_t n, const pg_rp *in, const int32_t *chr_a, const int32_t *chr_b, pg_rp_cluster *out)
{
   rpemu_modify_interchr(n, (const PgRp *)in, chr_a, chr_b, (PgRpCluster *)out);
   ctx->rp_launches++;
   return PG_OK;
}
int pg_rp_summarize_interchr(pg_ctx *ctx, int32_t n, const pg_rp *in, const int32_t *chr_a, const int32_t *chr_b, pg_rp_cluster *inout)
{
   rpemu_summarize_interchr(n, (const PgRp *)in, chr_a, chr_b, (PgRpCluster *)inout);
   ctx->rp_launches++;
   return PG_OK;
}
uint64_t pg_rp_launches(pg_ctx *ctx) { return ctx ? ctx->rp_launches : 0; }
int pg_box_sort(pg_ctx *ctx, uint32_t n_boxes, const uint32_t *box_off, const pg_sort_key *keys, uint32_t *order, uint8_t *nonunique)
{
   boxemu_sort(n_boxes, box_off, (const PgSortKey *)keys, order, nonunique);
   ctx->bx_launches++;
   return PG_OK;
}
int pg_box_sort_kind(pg_ctx *ctx, int32_t kind, uint32_t n_boxes, const uint32_t *box_off, const pg_sort_key *keys, uint32_t *order, uint8_t *nonunique, uint8_t *event)
{
   /* the kernel's formulation (chunked prefix minima, ballots written out), not the literal loops */
   boxemu_sort_warp_formulation_kind(kind, n_boxes, box_off, (const PgSortKey *)keys, order, nonunique, event);
   ctx->bx_launches++;
   return PG_OK;
}
uint64_t pg_box_sort_launches(pg_ctx *ctx) { return ctx ? ctx->bx_launches : 0; }
int64_t pg_bgzf_inflated_size(const uint8_t *file, size_t file_len) { return bgzfemu_size(file, file_len); }
int pg_bgzf_inflate(pg_ctx *ctx, const uint8_t *file, size_t file_len, uint8_t *out, size_t out_cap, size_t *out_len, int32_t check_crc)
{
   const long total = bgzfemu_size(file, file_len);
   if (total < 0) return fail(ctx, PG_ERR_DATA, "not a BGZF file");
   *out_len = (size_t)total;
   if (out_cap < (size_t)total) return fail(ctx, PG_ERR_RUNS_CAP, "output buffer too small");
   std::vector<uint32_t> status((size_t)(file_len / 26 + 2));
   const long nb = bgzfemu_inflate(file, file_len, out, status.data(), (long)status.size(), check_crc);
   for (long i = 0; i < nb; i++) if (status[(size_t)i]) return fail(ctx, PG_ERR_DATA, "corrupt block");
   ctx->stream.assign(out, out + total);      // "resident" for pg_bam_index
   return PG_OK;
}
int pg_bam_index(pg_ctx *ctx, pg_bam_rec *recs, uint64_t cap, uint64_t *n_recs, uint64_t *first_record, int32_t *n_ref)
{
   if (ctx->stream.empty()) return fail(ctx, PG_ERR_STATE, "no inflated stream");
   uint64_t stats[3] = {0, 0, 0};
   const long n = bamemu_index(ctx->stream.data(), ctx->stream.size(), 0, NULL, 0, stats);
   if (n < 0) return fail(ctx, PG_ERR_DATA, "not a BAM stream / corrupt chain");
   *n_recs = (uint64_t)n;
   if (first_record) *first_record = stats[2];
   if (n_ref) { int32_t r = 0; pgb_header(ctx->stream.data(), ctx->stream.size(), &r); *n_ref = r; }
   if ((uint64_t)n > cap || (n && !recs)) return fail(ctx, PG_ERR_RUNS_CAP, "record buffer too small");
   bamemu_index(ctx->stream.data(), ctx->stream.size(), 0, (PgBamRec *)recs, cap, stats);
   ctx->bam_table.assign((const PgBamRec *)recs, (const PgBamRec *)recs + n);
   return PG_OK;
}
int pg_bam_index_run(pg_ctx *ctx, uint64_t first_record, int32_t n_ref, pg_bam_rec *recs, uint64_t cap, uint64_t *n_recs, uint64_t *tail)
{
   if (ctx->stream.empty()) return fail(ctx, PG_ERR_STATE, "no inflated stream");
   if (first_record > ctx->stream.size()) return fail(ctx, PG_ERR_ARG, "first_record beyond the stream");
   uint64_t stats[4] = {0, 0, 0, 0};
   const long n = bamemu_index_run(ctx->stream.data(), ctx->stream.size(), first_record, n_ref, 0, NULL, 0, stats);
   if (n < 0) return fail(ctx, PG_ERR_DATA, "impossible block_size in the chain");
   *n_recs = (uint64_t)n;
   *tail = stats[3];
   if ((uint64_t)n > cap || (n && !recs)) return fail(ctx, PG_ERR_RUNS_CAP, "record buffer too small");
   bamemu_index_run(ctx->stream.data(), ctx->stream.size(), first_record, n_ref, 0, (PgBamRec *)recs, cap, stats);
   ctx->bam_table.assign((const PgBamRec *)recs, (const PgBamRec *)recs + n);
   return PG_OK;
}
int pg_bam_rp_discovery(pg_ctx *ctx, const pg_rp_query *q, pg_rp *rows, uint32_t *src, uint64_t cap, uint64_t *n_rows, pg_rp *inter_rows,
                        uint32_t *inter_src, uint64_t inter_cap, uint64_t *n_inter, uint64_t *n_scanned)
{
   if (ctx->stream.empty()) return fail(ctx, PG_ERR_STATE, "no indexed stream");
   const size_t n = ctx->bam_table.size();
   std::vector<PgRpRow> a(n + 1), b(n + 1);
   std::vector<uint32_t> sa(n + 1), sb(n + 1);
   uint64_t counts[3] = {0, 0, 0};
   bamemu_rp_discovery(ctx->stream.data(), ctx->bam_table.data(), (uint32_t)n, (const PgRpQuery *)q, a.data(), sa.data(), b.data(), sb.data(), counts);
   *n_rows = counts[0];
   *n_inter = counts[1];
   if (n_scanned) *n_scanned = counts[2];
   if (counts[0] > cap || counts[1] > inter_cap) return fail(ctx, PG_ERR_RUNS_CAP, "row buffers too small");
   if (counts[0]) { memcpy(rows, a.data(), sizeof(pg_rp) * counts[0]); memcpy(src, sa.data(), 4 * counts[0]); }
   if (counts[1]) { memcpy(inter_rows, b.data(), sizeof(pg_rp) * counts[1]); memcpy(inter_src, sb.data(), 4 * counts[1]); }
   return PG_OK;
}
int pg_bam_sr_candidates(pg_ctx *ctx, const pg_sr_query *q, pg_sr_cand *cands, uint64_t cand_cap, uint64_t *n_cands, char *seqs, uint64_t seq_cap,
                         uint32_t *stride, pg_sr_ref *refs, uint64_t ref_cap, uint64_t *n_refs, uint64_t *n_scanned)
{
   if (ctx->stream.empty()) return fail(ctx, PG_ERR_STATE, "no indexed stream");
   uint64_t counts[5] = {0, 0, 0, 0, 0};
   const uint32_t n = (uint32_t)ctx->bam_table.size();
   if (sremu_candidates(ctx->stream.data(), ctx->bam_table.data(), n, (const PgSrQuery *)q, 17u, 0, 0, NULL, NULL, NULL, counts) != 0)
      return fail(ctx, PG_ERR_DATA, "name table");
   if (n_scanned) *n_scanned = counts[3];
   if (counts[4] & PGSR_ERR_MALFORMED) return fail(ctx, PG_ERR_DATA, "an alignment record of the window is malformed");
   if (counts[4]) return fail(ctx, PG_ERR_DATA, "the insert size is not larger than a candidate read");
   const uint32_t str = counts[2] ? (uint32_t)counts[2] : 1u;
   *n_cands = counts[0]; *n_refs = counts[1]; *stride = str;
   if (counts[0] > cand_cap || counts[1] > ref_cap || counts[0] * str > seq_cap || (counts[0] && (!cands || !seqs)) || (counts[1] && !refs))
      return fail(ctx, PG_ERR_RUNS_CAP, "buffers too small");
   std::vector<uint8_t> tmp((size_t)counts[0] * (counts[2] ? counts[2] : 1) + 1, 0);
   sremu_candidates(ctx->stream.data(), ctx->bam_table.data(), n, (const PgSrQuery *)q, 17u, 0, 0, (PgSrCand *)cands, tmp.data(), (PgSrRef *)refs, counts);
   for (uint64_t k = 0; k < counts[0]; k++) memcpy(seqs + k * str, tmp.data() + k * counts[2], counts[2]);
   return PG_OK;
}
int pg_bam_sr_close_end(pg_ctx *ctx, const pg_sr_query *q, int32_t chr_id, pg_sr_cand *cands, uint64_t cand_cap, uint64_t *n_cands, char *seqs,
                        uint64_t seq_cap, uint32_t *stride, pg_sr_ref *refs, uint64_t ref_cap, uint64_t *n_refs, uint64_t *n_scanned, pg_results *res)
{
   /* the filter, then GetCloseEnd on its rows in their own layout: sequences at the stride, NUL-filled */
   std::vector<pg_sr_cand> c;
   std::vector<pg_sr_ref> rf;
   std::vector<char> sq;
   uint64_t n = 0, m = 0;
   uint32_t str = 0;
   int rc = pg_bam_sr_candidates(ctx, q, NULL, 0, &n, NULL, 0, &str, NULL, 0, &m, n_scanned);
   if (rc != PG_OK && rc != PG_ERR_RUNS_CAP) return rc;
   c.resize(n + 1); rf.resize(m + 1); sq.assign(n * str + 1, 0);
   if (rc == PG_ERR_RUNS_CAP) {
      rc = pg_bam_sr_candidates(ctx, q, c.data(), n, &n, sq.data(), n * str, &str, rf.data(), m, &m, n_scanned);
      if (rc != PG_OK) return rc;
   }
   *n_cands = n; *n_refs = m; *stride = str;
   std::vector<uint32_t> off(n + 1), pos(n);
   std::vector<char> md(n);
   std::vector<int16_t> isz(n, (int16_t)q->insert_size);
   std::vector<int32_t> chr(n, chr_id);
   for (uint64_t k = 0; k <= n; k++) off[k] = (uint32_t)(k * str);
   for (uint64_t k = 0; k < n; k++) { md[k] = (char)c[k].matched_d; pos[k] = c[k].matched_rel_pos; }
   std::vector<pg_read_out> out(n + 1);
   std::vector<pg_run> runs;
   uint32_t n_runs = 0;
   if (n) {
      pg_reads in;
      memset(&in, 0, sizeof in);
      in.n = (int32_t)n; in.seq_blob = sq.data(); in.seq_off = off.data(); in.matched_d = md.data(); in.matched_rel_pos = pos.data();
      in.insert_size = isz.data(); in.chr_id = chr.data();
      runs.resize(n * 16 + 1024);
      pg_results r;
      r.out = out.data(); r.runs = runs.data(); r.runs_cap = (uint32_t)runs.size(); r.n_runs = 0;
      rc = pg_close_end(ctx, &in, &r);
      if (rc == PG_ERR_RUNS_CAP) { runs.resize(r.n_runs); r.runs = runs.data(); r.runs_cap = r.n_runs; rc = pg_close_end(ctx, &in, &r); }
      if (rc != PG_OK) return rc;
      n_runs = r.n_runs;
   }
   res->n_runs = n_runs;
   if (n > cand_cap || m > ref_cap || n * str > seq_cap || (n && (!cands || !seqs || !res->out)) || (m && !refs) || n_runs > res->runs_cap ||
       (n_runs && !res->runs))
      return fail(ctx, PG_ERR_RUNS_CAP, "buffers too small");
   if (n) {
      memcpy(cands, c.data(), sizeof(pg_sr_cand) * n); memcpy(seqs, sq.data(), n * str);
      memcpy(res->out, out.data(), sizeof(pg_read_out) * n);
      if (n_runs) memcpy(res->runs, runs.data(), sizeof(pg_run) * n_runs);
   }
   if (m) memcpy(refs, rf.data(), sizeof(pg_sr_ref) * m);
   return PG_OK;
}
int pg_bd_load_region(pg_ctx *ctx, const pg_bd_event *events, uint32_t n_events, const int32_t *chr_name_rank, int32_t chr, uint32_t start, uint32_t end)
{
   ctx->bd_loaded = false;
   if (pgbd_prepare(reinterpret_cast<const PgBdEvent *>(events), n_events, chr_name_rank, (int32_t)ctx->chr.size(), chr, start, end, ctx->bd) != 0)
      return fail(ctx, PG_ERR_ARG, "events must be in the reference's order");
   ctx->bd_loaded = true;
   return PG_OK;
}
int pg_bd_unload(pg_ctx *ctx) { ctx->bd_loaded = false; return PG_OK; }
int pg_bd_clusters(pg_ctx *ctx, uint32_t n, const uint32_t *close_last_abs_loc, uint32_t *region_off, pg_region *regions, uint64_t cap, uint64_t *n_regions)
{
   if (!ctx->bd_loaded) return fail(ctx, PG_ERR_STATE, "no region loaded");
   PgBdRegionView R;
   R.start = ctx->bd.start.data(); R.drop = ctx->bd.drop.data(); R.sel = ctx->bd.sel.data(); R.m = (uint32_t)ctx->bd.start.size();
   R.z = ctx->bd.z; R.ws = ctx->bd.ws; R.we = ctx->bd.we;
   uint64_t tot = 0;
   for (uint32_t i = 0; i < n; i++) {
      uint32_t S, E;
      region_off[i] = (uint32_t)tot;
      pgbd_range(R, close_last_abs_loc[i], S, E);
      if (E > S) tot += pgbd_cluster(R, S, E, (regions && tot < cap) ? reinterpret_cast<PgRegion *>(regions + tot) : NULL, (uint32_t)(cap > tot ? cap - tot : 0));
   }
   if (n) region_off[n] = (uint32_t)tot;
   *n_regions = tot;
   ctx->bd_lookups += n;
   return (tot > cap || (tot && !regions)) ? fail(ctx, PG_ERR_RUNS_CAP, "region buffer too small") : PG_OK;
}
uint64_t pg_bd_launches(pg_ctx *) { return 0; }
uint64_t pg_bd_lookups(pg_ctx *ctx) { return ctx ? ctx->bd_lookups : 0; }
uint64_t pg_classify_launches(pg_ctx *ctx) { return ctx ? ctx->sv_launches : 0; }
double pg_classify_last_kernel_ms(pg_ctx *) { return 0.0; }
uint32_t pg_expand_runs(const pg_run *runs, uint32_t n_runs, pg_up *out) { return pgh_expand_runs(runs, n_runs, out); }
uint32_t pg_clean_unique_points(pg_up *ups, uint32_t n) { return pgh_clean_unique_points(ups, n); }
uint32_t pg_compress_ups(const pg_up *ups, uint32_t n, pg_run *out) { return pgh_compress_ups(ups, n, out); }
}
