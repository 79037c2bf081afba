/*
 * kmerjs_b200.h -- C ABI of libkmerjs_b200.so: the B200 (sm_100a) implementation of the
 * kmerjs hot path  FASTQ -> prefix-filtered k-mer counts -> KmerFinder template scoring.
 *
 * The reference (josl/kmerjs) has no FFI: the path sits behind three ES-module seams.  Each
 * group of entry points below replaces one of them (paths relative to the reference root):
 *
 *   kj_counts_*      KmerJS#readFile + kmersInLine + complement   lib/kmers.js:31-38,88-100,106-185
 *   kj_db_*          the k-mer -> template-list store              lib/kmerFinderServer.js:68-92,171-226,712-728
 *   kj_first_match   findFirstMatch result contract                lib/kmerFinderClient.js:128-173, lib/kmerFinderServer.js:171-226
 *   kj_wta_next      findMatches generator (winner takes all)      lib/kmerFinderClient.js:174-290
 *   kj_stats_*       zScore / fastp                                lib/stats.js:19-45,52-115
 *
 * Conventions: plain pointers and sizes only; all sizes uint64_t; caller owns every buffer it
 * passes in; the library owns every kj_* handle (freed only by the matching *_free/_destroy).
 * Every function returns KJ_OK (0) or a negative KJ_E_* code (kj_wta_next: 1 = row produced,
 * 0 = finished); kj_last_error() gives the text.  No exceptions cross the boundary.  There is
 * NO CPU fallback: without an sm_100 device kj_init fails with KJ_E_NO_SM100.
 * Thread model: callable from any host thread; one call at a time per kj_ctx (internal mutex).
 */
#ifndef KMERJS_B200_H
#define KMERJS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KJ_ABI_VERSION 1

enum {
    KJ_OK = 0,
    KJ_E_INVALID = -1,     /* bad argument */
    KJ_E_NO_SM100 = -2,    /* no Blackwell (sm_100) device: there is no CPU fallback */
    KJ_E_CUDA = -3,        /* CUDA runtime error (text in kj_last_error) */
    KJ_E_NOMEM = -4,
    KJ_E_IO = -5,
    KJ_E_TABLE_FULL = -6,  /* count table could not grow (raise capacity_hint) */
    KJ_E_NO_HITS = -7,     /* 'No hits were found!'  lib/kmerFinderServer.js:219-221, kmerFinderClient.js:264-266 */
    KJ_E_NO_WINNER = -8,   /* 'No hits were found! (kmerResults.length === 0)'  lib/kmerFinderClient.js:283-285 */
    KJ_E_RANGE = -9,       /* k > 32, line longer than the halo / position field, ... */
    KJ_E_STATE = -10       /* call order violated (e.g. add after finish) */
};

/* where a buffer lives */
enum { KJ_MEM_HOST = 0, KJ_MEM_DEVICE = 1 };

/* kj_count_params.flags */
enum {
    KJ_F_NO_ORDER = 1u << 0,     /* do not track first-seen ordinals (export order undefined) */
    KJ_F_FORCE_GENERIC = 1u << 1,/* use the line-oriented kernel even where the filter kernel applies (tests) */
    KJ_F_FORWARD_ONLY = 1u << 2, /* scan the line only, not complement(line): one KmerJS#kmersInLine call (lib/kmers.js:88-100) */
    KJ_F_NO_LINE_GATE = 1u << 3, /* drop readFile's `line.length > 1` gate (lib/kmers.js:151); kmersInLine has none */
    KJ_F_COUNT_BASES = 1u << 4   /* also sum the sequence-line lengths (kj_counts_bases); off by default: a statistic the
                                    reference does not have, and it costs the scan kernel about 20 % */
};

typedef struct kj_ctx kj_ctx;
typedef struct kj_counts kj_counts;
typedef struct kj_db kj_db;
typedef struct kj_match kj_match;

/* ---------------------------------------------------------------- context */

/* One context per process per GPU (one process per GPU; multi-GPU = torch.distributed/NCCL in the
 * host layer).  `stream` may be NULL (library creates its own) or a cudaStream_t to launch on. */
int kj_init(int device, void *stream, kj_ctx **out);
void kj_destroy(kj_ctx *ctx);
const char *kj_last_error(const kj_ctx *ctx); /* ctx may be NULL: last error of a failed kj_init */
int kj_abi_version(void);
/* number of kernels this library has launched on the context so far (bench.py: gpu_launches) */
uint64_t kj_launch_count(const kj_ctx *ctx);
/* average device time (ms, CUDA events on the launching stream) of the dominant extraction kernel
 * since the last kj_reset_timers(); n_launches receives the number of launches averaged */
double kj_scan_kernel_ms(const kj_ctx *ctx, uint64_t *n_launches);
/* the part of it spent in the candidate check + hash update kernel that follows the scan kernel */
double kj_verify_kernel_ms(const kj_ctx *ctx);
/* input bytes those launches owned (the F term of the algorithmic-bytes model) */
uint64_t kj_scan_kernel_bytes(const kj_ctx *ctx);
/* size of the chunks in which host buffers and files are staged into device memory (two pinned + two device buffers of
 * this size per context; default 64 MiB, or KJ_STAGE_CHUNK_MB from the environment at first use) */
int kj_set_stage_chunk(kj_ctx *ctx, uint64_t bytes);
void kj_reset_timers(kj_ctx *ctx);
void kj_enable_timers(kj_ctx *ctx, int on);

/* ---------------------------------------------------------------- extraction + count
 * replaces lib/kmers.js:88-100 (kmersInLine), :31-38 (complement), :106-185 (readFile). */

typedef struct kj_count_params {
    const uint8_t *prefix;   /* KmerJS `preffix` (lib/kmers.js:67,70); any bytes; may be empty */
    uint32_t prefix_len;
    uint32_t k;              /* `kmerLength` (1..32) */
    uint32_t step;           /* `step` (>=1); step>1 reproduces the short-window quirk of lib/kmers.js:89-99 */
    uint32_t flags;          /* KJ_F_* */
    uint64_t base_line;      /* number of '\n' in the stream before the first buffer (sharded ingest) */
    uint64_t base_col;       /* bytes of the current line that precede the first buffer */
    uint64_t capacity_hint;  /* expected distinct k-mers (0 = derive from the input size) */
} kj_count_params;

int kj_counts_create(kj_ctx *ctx, const kj_count_params *p, kj_counts **out);

/* Append the next piece of the byte stream.  `buf[0..n)` is readable; window starts in
 * `[0, own_n)` belong to this call, bytes `[own_n, n)` are a halo that the next call presents
 * again as its first bytes.  final != 0: `buf+n` is end of stream (then own_n must equal n);
 * otherwise n - own_n must be >= 32 (>= the longest line when step > 1 or the prefix is empty).
 * KJ_MEM_DEVICE buffers must be 16-byte aligned and stay valid until kj_counts_finish.
 * KJ_MEM_HOST buffers are staged through pinned memory in chunks with copy/compute overlap. */
int kj_counts_add_buffer(kj_counts *c, const uint8_t *buf, uint64_t n, uint64_t own_n,
                         int mem_kind, int final);
/* Whole file (KmerJS#readFile, env='node', lib/kmers.js:138-139): reader threads fill two pinned staging buffers of
 * kj_set_stage_chunk bytes while the other one is copied to the device and counted.  A file that starts with the gzip
 * magic (1f 8b; any number of members) is inflated by zlib on the reader thread into the same buffers -- an ingest format
 * the reference does not have; kj_counts_bytes_read then counts the inflated bytes.  KJ_E_IO on read / stream errors. */
int kj_counts_add_file(kj_counts *c, const char *path);
/* Wait for the device, build the compact (key,count,ordinal) list.  After this the handle is
 * read-only for add_*; merge/score/export are allowed. */
int kj_counts_finish(kj_counts *c);

uint64_t kj_counts_size(const kj_counts *c);        /* kmerMapSize  lib/kmers.js:177 */
uint64_t kj_counts_lines(const kj_counts *c);       /* KmerJS#lines lib/kmers.js:164-165 */
uint64_t kj_counts_bases(const kj_counts *c);       /* sum of the lengths of the sequence lines (line index 1 mod 4); needs KJ_F_COUNT_BASES */
uint64_t kj_counts_bytes_read(const kj_counts *c);  /* KmerJS#bytesRead lib/kmers.js:146 */
uint64_t kj_counts_occurrences(const kj_counts *c); /* sum of all counts */
/* Export in first-insertion order (the order of JS Map iteration / mapToJSON, lib/kmers.js:46-54).
 * keys: size() * 32 bytes, key i at keys + 32*i, key_len[i] valid bytes (ASCII, not NUL padded). */
int kj_counts_export(kj_counts *c, uint8_t *keys, uint32_t *key_len, uint64_t *counts);
/* alive[i] (export order) = 1 while key i is still in the map: kj_wta_next deletes the winner's
 * k-mers from the query exactly as removeWinnerKmers does (lib/kmerFinderClient.js:220-230) */
int kj_counts_alive(kj_counts *c, uint8_t *alive);
void kj_counts_free(kj_counts *c);

/* sharded ingest: number of '\n' in a DEVICE buffer and offset + 1 of the last one (0 = none); a
 * rank needs the number of lines before its byte range because the record FSM of
 * lib/kmers.js:151-163 is "line index mod 4" */
int kj_count_newlines(kj_ctx *ctx, const uint8_t *buf, uint64_t n, int mem_kind,
                      uint64_t *n_newlines, uint64_t *last_newline_plus1);

/* multi-GPU exchange: packed records {u64 key2bit, u64 count, u64 ordinal} of the
 * regular (ACGT-only, full-length) k-mers, grouped by owner = kj_owner(key) % n_parts.
 * `*dev_records` stays valid until the next partition call or kj_counts_free. */
int kj_counts_partition(kj_counts *c, uint32_t n_parts, const void **dev_records,
                        uint64_t *part_sizes /* n_parts entries */);
/* fold packed records (device memory) into this table: counts add, ordinals min */
int kj_counts_merge_records(kj_counts *c, const void *dev_records, uint64_t n);
/* the same for records in host memory (a k-mer map that did not come from kj_counts_add_*: JSON) */
int kj_counts_merge_host_records(kj_counts *c, const void *host_records, uint64_t n);
/* irregular (non-ACGT / short) k-mers travel as 56-byte host records {u8 key[32], u64 len, u64 count, u64 ordinal} */
uint64_t kj_counts_irregular_size(const kj_counts *c);
int kj_counts_irregular_export(kj_counts *c, void *host_records);
int kj_counts_irregular_merge(kj_counts *c, const void *host_records, uint64_t n);
/* the same, keeping only the records this part owns (owner = hash of the padded key bytes and the length, the
 * rule kj_owner and kj_db_desc.part/n_parts use for byte-string k-mers): every rank can be handed all records */
int kj_counts_irregular_merge_part(kj_counts *c, const void *host_records, uint64_t n, uint32_t part, uint32_t n_parts);
/* Fixed-capacity exchange: the same redistribution without a size round trip and without a host wait.  The sender
 * scatters its table (as it stands after the adds: no kj_counts_finish needed) by owner into n_parts segments of
 * kj_segment_bytes(cap_reg, cap_irr) bytes each -- {header: record counts and the sender's line / base / occurrence /
 * byte totals | regular records[cap_reg] | irregular records[cap_irr]} -- the ranks exchange them with ONE equal-split
 * all-to-all, and the owner merges the n_parts segments it received; record counts are read on the device.  A
 * segment that would have needed more than its capacity, or a sender whose count was not complete, makes the
 * owner's kj_counts_finish fail with KJ_E_RANGE: the caller then falls back to kj_counts_partition.  After
 * kj_counts_merge_segments + kj_counts_finish the handle's lines / bases / occurrences / bytes_read are the job's. */
uint64_t kj_segment_bytes(uint32_t cap_reg, uint32_t cap_irr);
int kj_counts_partition_segments(kj_counts *c, uint32_t n_parts, void *dev_segments, uint32_t cap_reg, uint32_t cap_irr);
int kj_counts_merge_segments(kj_counts *c, const void *dev_segments, uint32_t n_parts, uint32_t cap_reg, uint32_t cap_irr);
/* totals of the whole job for a handle that holds only the k-mers one rank owns */
int kj_counts_set_totals(kj_counts *c, uint64_t lines, uint64_t bases, uint64_t occurrences,
                         uint64_t bytes_read);
/* owner of a k-mer given as ASCII (the same functions the device uses): full-length ACGT-only k-mers by their
 * 2-bit key, every other k-mer (N, lower case, ...: the byte-string side table) by a hash of its padded bytes */
uint32_t kj_owner(const uint8_t *kmer, uint32_t len, uint32_t n_parts);

/* ---------------------------------------------------------------- template database
 * k-mer -> ordered template list, per-template attributes, Summary record
 * (lib/kmerFinderServer.js:68-92,184-199,716-724; src/kmerPyToMongo.py:35-42). */

typedef struct kj_db_desc {
    uint64_t n_kmers;
    const uint8_t *kmer_bytes;   /* concatenated ASCII k-mers */
    const uint32_t *kmer_len;    /* n_kmers */
    const uint64_t *list_off;    /* n_kmers + 1 offsets into tmpl_ids */
    const uint32_t *tmpl_ids;    /* template ids in DB list order */
    uint32_t n_templates;
    const uint64_t *lengths;     /* per template `lengths` */
    const uint64_t *ulengths;    /* per template `ulengths` */
    uint64_t summary_templates;  /* Summary.templates */
    uint64_t summary_unique_lens;/* Summary.uniqueLens */
    uint64_t summary_total_len;  /* Summary.totalLen */
    uint32_t part, n_parts;      /* keep only the k-mers with kj_owner(kmer) == part (n_parts 0/1: all) */
} kj_db_desc;

int kj_db_create(kj_ctx *ctx, const kj_db_desc *d, kj_db **out);
/* From disk (the step before the path): the reference's own JSON layouts and a versioned packed binary.
 *   KJ_DB_KMER_DOCS       [{"kmer", "templates": [{"sequence","lengths","ulengths","species"}]}]  lib/kmerFinderServer.js:68-92,184-199
 *   KJ_DB_TEMPLATE_DOCS   [{"sequence","lengths","ulenght","species","reads":[..]}]                src/kmerPyToMongo.py:35-42
 *   KJ_DB_KMERFINDER_MAP  {kmer: "T1,T2,.."} + <path>.lengths.json / .ulengths.json / .descriptions.json   lib/index.js:184-192
 *   KJ_DB_PACKED          written by kj_db_save_packed
 *   KJ_DB_AUTO            decided from the content
 * summary_path: a Summary record {"templates","uniqueLens","totalLen"} (lib/kmerFinderServer.js:716-724); NULL derives it
 * from the templates (packed files carry their own).  part / n_parts as in kj_db_desc. */
enum { KJ_DB_AUTO = 0, KJ_DB_KMER_DOCS = 1, KJ_DB_TEMPLATE_DOCS = 2, KJ_DB_KMERFINDER_MAP = 3, KJ_DB_PACKED = 4 };
int kj_db_load(kj_ctx *ctx, const char *path, int format, const char *summary_path, uint32_t part, uint32_t n_parts,
               kj_db **out);
/* write a host description as the packed binary (names / species: n_templates C strings, may be NULL); needs no device */
int kj_db_save_packed(const char *path, const kj_db_desc *d, const char *const *names, const char *const *species);
/* template attributes (name / species are empty strings unless the database came through kj_db_load) and the Summary */
int kj_db_template(const kj_db *db, uint32_t id, const char **name, const char **species, uint64_t *lengths,
                   uint64_t *ulength);
int kj_db_summary(const kj_db *db, uint64_t *templates, uint64_t *unique_lens, uint64_t *total_len);
void kj_db_free(kj_db *db);
uint64_t kj_db_n_kmers(const kj_db *db);
uint64_t kj_db_n_pairs(const kj_db *db);            /* (k-mer, template) pairs held */
uint32_t kj_db_n_templates(const kj_db *db);

/* ---------------------------------------------------------------- scoring */

/* findFirstMatch: per-template uScore/tScore over the query, hits = sum of list lengths.
 * The counts handle must be finished; it is consumed read-only except for the alive mask that
 * kj_wta_next maintains (mirrors kmerMap.delete, lib/kmerFinderClient.js:220-230). */
int kj_first_match(kj_ctx *ctx, kj_counts *q, const kj_db *db, kj_match **out);
uint64_t kj_match_hits(const kj_match *m);
uint32_t kj_match_n_matched(const kj_match *m);   /* templates with uScore > 0 */
/* uscore/tscore: n_templates entries each (0 for unmatched); order: ids of the matched templates
 * in first-encounter order (lib/kmerFinderServer.js:180-201), n_matched entries */
int kj_match_scores(kj_match *m, uint64_t *uscore, uint64_t *tscore, uint32_t *order);
/* the `kmers` Set of a template in the findFirstMatch reply (lib/kmerFinderClient.js:150-157): positions, in the
 * export order of the counts handle, of the query k-mers that list the template; ascending = insertion order.
 * idx may be NULL to ask for the count. */
int kj_match_template_kmers(kj_match *m, uint32_t template_id, uint64_t *idx, uint64_t cap, uint64_t *n);
void kj_match_free(kj_match *m);

/* multi-GPU protocol (SURVEY 8e collectives 3/4).  The query is sharded by k-mer owner and so is
 * the DB; every rank holds partial sums.  The host layer reduces them over NCCL in buffers it owns
 * (torch tensors): kj_match_get copies a vector into a caller-provided DEVICE buffer, kj_match_set
 * takes the reduced values back.
 *   m = kj_first_match_local(q, db)
 *   get/allreduce(SUM)/set  KJ_VEC_SCORES      u64[2T+1] = {uScore[T], tScore[T], hits}
 *   get/allreduce(MIN)/set  KJ_VEC_FIRST_ORD   u64[T]  first-seen ordinal of the first matching k-mer
 *   get/allreduce(MIN)/set  KJ_VEC_FIRST_IDX   u64[T]  its DB list index (computed against the reduced ordinals)
 *   kj_match_set_query_size(global kmerMapSize); kj_match_commit(m)
 *   per round: kj_wta_next (argmax on the global sums, removal on the local shard), then
 *              get/allreduce(SUM)/set KJ_VEC_SCORES again.
 * kj_first_match = kj_first_match_local + kj_match_commit + the 'No hits were found!' check. */
enum { KJ_VEC_SCORES = 0, KJ_VEC_FIRST_ORD = 1, KJ_VEC_FIRST_IDX = 2 };
int kj_first_match_local(kj_ctx *ctx, kj_counts *q, const kj_db *db, kj_match **out);
uint64_t kj_match_vec_len(kj_match *m, int which);          /* in u64 elements */
int kj_match_get(kj_match *m, int which, void *dev_out);
int kj_match_set(kj_match *m, int which, const void *dev_in);
/* the same without waiting for the copy: stream-ordered on the context's stream.  For callers whose
 * collective runs on (or is ordered against) that stream -- kj_init(device, stream) with the stream the
 * NCCL calls are issued from -- so that a WTA round costs one host synchronisation instead of four */
int kj_match_get_async(kj_match *m, int which, void *dev_out);
int kj_match_set_async(kj_match *m, int which, const void *dev_in);
int kj_match_commit(kj_match *m);
/* override kmerMapSize (the global query size when the query is sharded over ranks) */
int kj_match_set_query_size(kj_match *m, uint64_t kmer_map_size);

typedef struct kj_row {       /* lib/kmerFinderClient.js:75-89, same order */
    uint32_t template_id;     /* 'template' / 'species' are resolved by the host layer */
    uint32_t reserved;
    uint64_t score;           /* uScore of the winner this round */
    double expected;
    double z;                 /* rounded to 2 dp with the context rounding mode */
    double probability;
    double frac_q;
    double frac_d;
    double depth;
    uint64_t kmers_template;  /* ulength */
    double total_frac_q;
    double total_frac_d;
    double total_temp_cover;
    /* extras (not part of the reference row) */
    uint64_t tscore;
    uint64_t hits;            /* results.hits of this round */
    double z_device;          /* double-precision z computed on the device (unrounded) */
    double probability_device;
} kj_row;

/* One step of the findMatches generator: 1 = row written, 0 = loop ended normally,
 * KJ_E_NO_HITS / KJ_E_NO_WINNER mirror the two throws of lib/kmerFinderClient.js:264-266,283-285. */
int kj_wta_next(kj_match *m, kj_row *out);
/* Deferred rows, for callers that issue a collective per round (multi-GPU): with kj_match_defer_rows(m, 1),
 * kj_wta_next may return 2 = "winner accepted by the device gate (away from every fastp threshold), its k-mers
 * are being removed, integers in out->score/tscore/hits"; the caller starts its all-reduce and then calls
 * kj_wta_row, which finishes the row in exact-decimal arithmetic while the GPU works. */
int kj_match_defer_rows(kj_match *m, int on);
int kj_wta_row(kj_match *m, kj_row *out);
/* The whole generator in one call (no trip through the host language per row).  rows: capacity cap (>= maxHits); *n_rows rows were yielded; *end_status = 0 (normal end)
 * or the KJ_E_NO_HITS / KJ_E_NO_WINNER the generator throws after them (text in kj_last_error). */
int kj_wta_all(kj_match *m, kj_row *rows, uint32_t cap, uint32_t *n_rows, int *end_status);
/* Gathered matches (multi-GPU): instead of a collective per winner-takes-all round, every rank exports the
 * query entries that hit its DB shard as self-contained records, the ranks all-gather them (NVLink bandwidth,
 * one exchange), and each rank runs the loop of lib/kmerFinderClient.js:233-289 on the whole matched set.
 *   kj_match_matched_size   : entries that hit and the sum of their template-list lengths (synchronises);
 *   kj_match_export_matched : device buffers, entries = u64[4] {count, ordinal, list begin, list length}
 *                             per entry (begin relative to dev_tmpl), dev_tmpl = u32 template ids in DB order;
 *                             stream-ordered on the context's stream;
 *   kj_match_from_matched   : a match over the segments of all ranks (device addresses as u64), template
 *                             metadata from db (any shard: lengths / Summary are replicated).  Follow with
 *                             kj_match_commit.  kj_match_template_kmers is not available on it. */
int kj_match_matched_size(kj_match *m, uint64_t *n_entries, uint64_t *n_pairs);
int kj_match_export_matched(kj_match *m, void *dev_entries, uint64_t cap_entries, void *dev_tmpl, uint64_t cap_pairs);
int kj_match_from_matched(kj_ctx *ctx, const kj_db *db, uint32_t n_segments, const uint64_t *seg_entries,
                          const uint64_t *seg_n_entries, const uint64_t *seg_tmpl, const uint64_t *seg_n_pairs,
                          uint64_t kmer_map_size, kj_match **out);
/* The same with fixed capacities (no size round trip, no host wait before the collective): every rank writes ONE
 * segment of kj_matched_segment_bytes(cap_entries, cap_pairs) bytes -- {entries, pairs, query size, flags | entries |
 * template ids} --, the ranks all-gather the segments, and kj_match_from_segments builds the match over the gathered
 * buffer (which must stay valid until kj_match_free: the template lists are used in place).  Sizes are read on the
 * device; kj_match_commit fails with KJ_E_RANGE when a segment overflowed or a rank passed flags != 0, and sets the
 * query size to the sum of the ranks' (kj_match_query_size). */
uint64_t kj_matched_segment_bytes(uint32_t cap_entries, uint32_t cap_pairs);
int kj_match_export_segment(kj_match *m, void *dev_segment, uint32_t cap_entries, uint32_t cap_pairs,
                            uint64_t query_size, uint64_t flags);
/* The segment of a rank straight from a handle's hash table, WITHOUT kj_counts_finish in between (the owner side of
 * kj_counts_merge_segments: its finish then leaves the chain count -> exchange -> gather -> winner-takes-all).  Query size
 * and the "this rank's exchange did not fit" flag come from the handle's device counters.  Returns 1 -- not an error --
 * when the short cut does not apply (the DB holds byte-string k-mers or the all-G 32-mer, k differs, a piece is still in
 * flight): finish the handle and use kj_first_match_local + kj_match_export_segment.
 * Replaces nothing in the reference (lib/kmerFinderServer.js:171-226 queries one Redis for the whole map). */
int kj_counts_export_matched_segment(kj_counts *c, const kj_db *db, void *dev_segment, uint32_t cap_entries,
                                     uint32_t cap_pairs);
int kj_match_from_segments(kj_ctx *ctx, const kj_db *db, uint32_t n_segments, const void *dev_segments,
                           uint32_t cap_entries, uint32_t cap_pairs, kj_match **out);
uint64_t kj_match_query_size(const kj_match *m);
int kj_match_segment_sizes(const kj_match *m, uint64_t *n_entries, uint64_t *n_pairs);   /* summed over the ranks */
/* maxHits (lib/kmerFinderClient.js:123), default 100 */
int kj_match_set_max_hits(kj_match *m, uint32_t max_hits);
/* standardScoring (lib/kmerFinderServer.js:857-874): one row per matched template from the first
 * match, sorted by score descending (stable); rows the evalue gate rejects are skipped.
 * rows: capacity n_rows_cap; *n_rows receives the count. */
int kj_standard_scoring(kj_match *m, kj_row *rows, uint32_t n_rows_cap, uint32_t *n_rows);

/* ---------------------------------------------------------------- stats (lib/stats.js) */

/* bignumber.js ROUNDING_MODE used by dividedBy/sqrt/round(dp): 4 = HALF_UP (default),
 * 2 = CEIL (what lib/kmerFinderServer.js:7 configures) */
int kj_set_rounding_mode(kj_ctx *ctx, int mode);
/* exact-decimal zScore (20 dp, as bignumber.js); returns the double nearest to the decimal.
 * Host-only (no GPU needed).  z_text (may be NULL) receives the decimal string. */
int kj_stats_zscore(int rounding_mode, uint64_t r1, uint64_t n1, uint64_t r2, uint64_t n2,
                    double *z, char *z_text, uint64_t z_text_cap);
/* fastp of a decimal string (exact compare) */
int kj_stats_fastp_text(const char *z_text, double *p);
/* the device implementation (double precision), run on the GPU for `n` tuples */
int kj_stats_zscore_device(kj_ctx *ctx, uint64_t n, const uint64_t *r1, const uint64_t *n1,
                           const uint64_t *r2, const uint64_t *n2, double *z, double *p);
/* exact-decimal row for given integers (host only): the function kj_wta_next uses to finish a row */
int kj_stats_row(int rounding_mode, uint64_t uscore, uint64_t tscore, uint64_t uscore0,
                 uint64_t tscore0, uint64_t lengths, uint64_t ulength, uint64_t hits,
                 uint64_t kmer_map_size, uint64_t summary_templates, uint64_t summary_unique_lens,
                 kj_row *out, int *accepted);

/* ---------------------------------------------------------------- synthetic reads (bench/test utility)
 * Deterministic Illumina-shaped FASTQ generated on the device.  Not part of the
 * reference path; lives here so bench.py can build inputs larger than host generation allows. */
typedef struct kj_synth_params {
    uint64_t seed;
    uint64_t n_reads;
    uint32_t read_len;         /* 150 */
    uint64_t first_read;       /* global index of the first read (sharding) */
    const uint8_t *genome;     /* DEVICE pointer, ASCII ACGT */
    uint64_t genome_len;
    double sub_rate;           /* substitution error per base */
    double n_rate;             /* per-base N */
    double lead_n_rate;        /* fraction of reads whose first base is N */
} kj_synth_params;
/* bytes needed for the records of p (exact) */
int kj_synth_size(kj_ctx *ctx, const kj_synth_params *p, uint64_t *n_bytes);
/* write the records to dev_out (device, >= n_bytes) */
int kj_synth_generate(kj_ctx *ctx, const kj_synth_params *p, uint8_t *dev_out, uint64_t n_bytes);
/* n uniform random ACGT bytes into dev_out (device) */
int kj_synth_genome(kj_ctx *ctx, uint64_t seed, uint8_t *dev_out, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* KMERJS_B200_H */
