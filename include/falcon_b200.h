/*
 * falcon_b200.h -- C ABI of libfalcon_b200.so, the B200-native fc_consensus engine.
 *
 * Drop-in boundary.  Two groups of entry points:
 *
 *  (1) LEGACY symbols -- exactly what falcon_kit/falcon_kit.py binds from the reference's
 *      falcon.so (reference: falcon_kit/falcon_kit.py:54-122, src/c/common.h:59-177).  Same names,
 *      same struct layouts, same ownership rules (callee allocates with malloc/calloc, caller
 *      releases with the matching free_* function).  generate_consensus() and align() run on the
 *      GPU (batch of one); the k-mer helper symbols operate on caller-visible host structures
 *      whose layout is part of the ABI and are host code.
 *
 *  (2) BATCHED symbols (fcx_*) -- the throughput path: a resident 2-bit read pool plus seed
 *      blocks that index into it; one call runs k-mer range finding, the banded O(ND) DP, the
 *      traceback and the alignment-graph consensus for every block on the device.
 *
 * No torch types, no C++ types: plain pointers and sizes.  All functions are synchronous with
 * respect to the host unless stated otherwise.  Errors: fcx_* return 0 on success, non-zero on
 * failure with a message in fcx_last_error(); the legacy symbols have no error channel in the
 * reference (it abort()s, DW_banded.c:100-113, falcon.c:343,476) and abort() here as well.
 */
#ifndef FALCON_B200_H
#define FALCON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- legacy types (common.h) */
typedef int seq_coor_t;                                   /* common.h:57 */

typedef struct {                                          /* common.h:59-69 */
    seq_coor_t aln_str_size;
    seq_coor_t dist;
    seq_coor_t aln_q_s;
    seq_coor_t aln_q_e;
    seq_coor_t aln_t_s;
    seq_coor_t aln_t_e;
    char *q_aln_str;
    char *t_aln_str;
} alignment;

typedef struct {                                          /* common.h:95-99 */
    seq_coor_t start;
    seq_coor_t last;
    seq_coor_t count;
} kmer_lookup;

typedef unsigned char base;                               /* common.h:101-104 */
typedef base *seq_array;
typedef seq_coor_t seq_addr;
typedef seq_addr *seq_addr_array;

typedef struct {                                          /* common.h:107-111 */
    seq_coor_t count;
    seq_coor_t *query_pos;
    seq_coor_t *target_pos;
} kmer_match;

typedef struct {                                          /* common.h:114-120 */
    seq_coor_t s1, e1, s2, e2;
    long int score;
} aln_range;

typedef struct {                                          /* common.h:123-126 */
    char *sequence;
    int *eqv;
} consensus_data;

/* ---------------------------------------------------------------- legacy entry points */
/* replaces src/c/falcon.c:562-666; bound at falcon_kit/mains/consensus.py:20-23 */
consensus_data *generate_consensus(char **input_seq, unsigned int n_seq, unsigned min_cov,
                                   unsigned K, double min_idt);
/* replaces src/c/falcon.c:776-780 */
void free_consensus_data(consensus_data *);

/* replaces src/c/DW_banded.c:115-330; bound at falcon_kit/falcon_kit.py:111-114 */
alignment *align(char *query_seq, seq_coor_t q_len, char *target_seq, seq_coor_t t_len,
                 seq_coor_t band_tolerance, int get_aln_str);
void free_alignment(alignment *);                         /* DW_banded.c:333-337 */

/* k-mer helpers, replace src/c/kmer_lookup.c:71-119,140-204,207-292,294-427,429-589;
 * bound at falcon_kit/falcon_kit.py:54-83 */
kmer_lookup *allocate_kmer_lookup(seq_coor_t size);
void init_kmer_lookup(kmer_lookup *, seq_coor_t size);
void free_kmer_lookup(kmer_lookup *);
seq_array allocate_seq(seq_coor_t size);
void init_seq_array(seq_array, seq_coor_t size);
void free_seq_array(seq_array);
seq_addr_array allocate_seq_addr(seq_coor_t size);
void free_seq_addr_array(seq_addr_array);
void add_sequence(seq_coor_t start, unsigned int K, char *seq, seq_coor_t seq_len,
                  seq_addr_array sda, seq_array sa, kmer_lookup *lk);
void mask_k_mer(seq_coor_t size, kmer_lookup *kl, seq_coor_t threshold);
kmer_match *find_kmer_pos_for_seq(char *seq, seq_coor_t seq_len, unsigned int K,
                                  seq_addr_array sda, kmer_lookup *lk);
void free_kmer_match(kmer_match *);
aln_range *find_best_aln_range(kmer_match *, seq_coor_t K, seq_coor_t bin_size,
                               seq_coor_t count_th);
aln_range *find_best_aln_range2(kmer_match *, seq_coor_t K, seq_coor_t bin_width,
                                seq_coor_t count_th);
void free_aln_range(aln_range *);

/* ---------------------------------------------------------------- batched GPU path */
typedef struct fcx_ctx fcx_ctx;

/* Number of CUDA devices visible to this process (0 if there is none: falcon_b200 has no CPU path). */
int fcx_device_count(void);

/* Create an engine bound to CUDA device `device` (one engine per process per GPU). */
int fcx_create(int device, fcx_ctx **out);
void fcx_destroy(fcx_ctx *);
/* Last error text for this engine (ctx may be NULL for creation errors). */
const char *fcx_last_error(const fcx_ctx *);

/* Pinned host staging memory for the caller's read bytes (optional; plain memory also works). */
void *fcx_host_alloc(size_t bytes);
void fcx_host_free(void *);

/* Upload a read pool: n_reads upper-case ACGT sequences stored back to back in `bases`
 * (no terminators needed), read r = bases[offsets[r] .. offsets[r+1]).  The pool is 2-bit packed
 * on the device and stays resident until the next fcx_pool_upload / fcx_destroy.  Any byte
 * outside "ACGT" is an error (the reference's behaviour is undefined for it: falcon.c:370-379
 * indexes base[-1]). */
int fcx_pool_upload(fcx_ctx *, const char *bases, const uint64_t *offsets, uint32_t n_reads);

/* The same in three steps, for pools assembled from parts (several processes / devices each
 * holding some of the reads; SURVEY.md 8(e): "one broadcast of the read index"):
 *   reserve      fixes the layout from the read lengths (offsets as above; no data is read) and
 *                returns the size of the packed pool in 32-bit words;
 *   upload_part  uploads and packs reads [first_read, first_read + n_part) -- `bases`/`offsets`
 *                describe just that part;
 *   device       exposes the packed pool (device pointer, word count, host array of n_reads + 1
 *                word offsets) so that the caller can fill the parts it did not upload with a
 *                collective (NCCL broadcast) or a peer copy straight into place;
 *   commit       makes the pool usable by fcx_consensus_blocks. */
int fcx_pool_reserve(fcx_ctx *, const uint64_t *offsets, uint32_t n_reads, uint64_t *total_words);
int fcx_pool_upload_part(fcx_ctx *, const char *bases, const uint64_t *offsets, uint32_t first_read,
                         uint32_t n_part);
int fcx_pool_device(fcx_ctx *, void **dev_words, uint64_t *n_words, const uint64_t **word_off);
int fcx_pool_commit(fcx_ctx *);

/* Consensus for n_blocks seed blocks.  Block b consists of pool reads
 * read_ids[block_off[b] .. block_off[b+1]); the first is the seed (target), the rest are aligned
 * to it in order (exactly the char** order of the reference's generate_consensus).
 * K must be 8 (falcon_kit/mains/consensus.py:270).
 * Results: *out_bases is one buffer holding the consensus strings back to back,
 * block b = [(*out_off)[b], (*out_off)[b+1]); both buffers are owned by the engine and stay
 * valid until the next fcx_consensus_blocks / fcx_destroy on this engine. */
int fcx_consensus_blocks(fcx_ctx *, uint32_t n_blocks, const uint32_t *block_off,
                         const uint32_t *read_ids, unsigned min_cov, unsigned K, double min_idt,
                         const char **out_bases, const uint64_t **out_off);

/* Per-pair diagnostics of the last fcx_consensus_blocks call (pair p = the p-th non-seed read in
 * block order).  Used by the parity tests to compare stage by stage with the oracle. */
typedef struct {
    int32_t n_match;               /* k-mer hits */
    int32_t s1, e1, s2, e2;        /* chosen ranges */
    int32_t passed_filter;         /* falcon.c:613-619 */
    int32_t aligned;               /* DP reached an end */
    int32_t dist;                  /* D */
    int32_t aln_size;              /* A */
    int32_t q_e, t_e;
    int32_t accepted;              /* falcon.c:629 */
    int32_t n_tags;                /* alignment columns voted */
    int32_t trace_cells;           /* E */
} fcx_pair_info;
int fcx_last_pair_info(fcx_ctx *, fcx_pair_info *out, uint64_t max_pairs, uint64_t *n_pairs);

/* Device-side timings (CUDA events on the engine's stream) of the last fcx_consensus_blocks call,
 * in milliseconds, and work counters.  Index with the FCX_T_* / FCX_C_* constants. */
enum { FCX_T_INDEX = 0, FCX_T_RANGE, FCX_T_DP, FCX_T_TRACEBACK, FCX_T_CONSENSUS, FCX_T_TOTAL,
       FCX_T_COUNT };
enum { FCX_C_PAIRS = 0, FCX_C_DP_PAIRS, FCX_C_ACCEPTED, FCX_C_TRACE_CELLS, FCX_C_DP_STEPS,
       FCX_C_ALN_COLS, FCX_C_SPAN_BASES, FCX_C_KERNEL_LAUNCHES, FCX_C_WAVES, FCX_C_COUNT };
int fcx_last_stats(fcx_ctx *, double *times_ms /*FCX_T_COUNT*/, uint64_t *counters /*FCX_C_COUNT*/);

/* ---------------------------------------------------------------- LA4Falcon stream parser
 * Host-side parser of the stdin block text (replaces the per-line Python of
 * falcon_kit/mains/consensus.py:161-209 and get_longest_reads :26-45; same rules, see
 * fcx_parser.cu).  feed() takes arbitrary chunks of the stream and returns the number of complete
 * blocks queued, or -(queued + 1) once the "- -" terminator has been seen; take() hands out up to
 * max_blocks blocks (and at most max_bases bases, but at least one block) as a pool + block lists
 * in exactly the shape fcx_pool_upload / fcx_consensus_blocks accept; seed_ids are NUL-separated.
 * The returned buffers are owned by the parser and stay valid until the SECOND next take() (two
 * sets are used alternately, so one thread can parse the next batch while another consumes this one). */
typedef struct fcx_parser fcx_parser;
fcx_parser *fcx_parser_create(unsigned min_n_read, unsigned min_len_aln, unsigned max_n_read,
                              unsigned min_cov_aln, unsigned max_cov_aln);
void fcx_parser_destroy(fcx_parser *);
int fcx_parser_feed(fcx_parser *, const char *data, size_t n, int eof);
int fcx_parser_pending(const fcx_parser *);
int fcx_parser_stopped(const fcx_parser *);
int fcx_parser_take(fcx_parser *, uint32_t max_blocks, uint64_t max_bases, const char **bases,
                    const uint64_t **offsets, uint32_t *n_reads, const uint32_t **block_off,
                    const uint32_t **read_ids, uint32_t *n_blocks, const char **seed_ids);

/* ---------------------------------------------------------------- Dazzler DB + .las input
 * Reads a Dazzler read database (<root>.db stub + hidden .<root>.idx / .<root>.bps) and a local
 * alignment file (.las) directly, replacing the text hop `LA4Falcon -H$CUTOFF -fo db las | consensus`
 * (falcon_kit/mains/consensus_task.py:81-90, falcon_kit/bash.py:349-358):
 *   fcx_dazz_upload  puts every (trimmed) read of the DB into the engine's pool, 2-bit packed, in both
 *                    orientations: pool id 2r = read r, 2r + 1 = its reverse complement;
 *   fcx_las_take     walks the overlap records with LA4Falcon's -f -o -H<seed_cutoff> rules and the
 *                    consensus parser's rules (falcon_kit/mains/consensus.py:161-209, :26-45) and
 *                    returns seed blocks as lists of pool ids (seed ids formatted %08d as LA4Falcon
 *                    prints them, NUL-separated), ready for fcx_consensus_blocks.
 * File layouts restate DAZZ_DB's DB.h and DALIGNER's align.h; see fcx_dazz.cu for provenance. */
typedef struct fcx_dazz fcx_dazz;
int fcx_dazz_open(const char *db_path, fcx_dazz **out);
void fcx_dazz_close(fcx_dazz *);
const char *fcx_dazz_last_error(const fcx_dazz *);
uint32_t fcx_dazz_nreads(const fcx_dazz *);
int32_t fcx_dazz_read_length(const fcx_dazz *, uint32_t read);
int fcx_dazz_upload(fcx_dazz *, fcx_ctx *);
int fcx_las_open(fcx_dazz *, const char *las_path);
int fcx_las_take(fcx_dazz *, int seed_cutoff, unsigned min_n_read, unsigned min_len_aln, unsigned max_n_read,
                 unsigned min_cov_aln, unsigned max_cov_aln, uint32_t max_blocks, uint64_t max_pairs,
                 const uint32_t **block_off, const uint32_t **read_ids, uint32_t *n_blocks,
                 const char **seed_ids, int *done);
/* Pool from a .bps image: n_reads reads of rlen[r] bases whose compressed bases start at byte boff[r];
 * 2 * n_reads pool entries (forward, reverse complement), each cut like consensus.py:178-179. */
int fcx_pool_upload_bps(fcx_ctx *, const uint8_t *bps, uint64_t n_bytes, const uint64_t *boff,
                        const int32_t *rlen, uint32_t n_reads);

/* Engine options (name, value):
 *   "pair_info"        0/1  keep the per-pair diagnostics of fcx_last_pair_info (default 1)
 *   "eqv"              0/1  also produce the eqv array of consensus_data (legacy symbol; default 0)
 *   "arena_gb"         device-memory budget for the wave buffers, GB (default 85 % of free memory)
 *   "max_wave_blocks" / "min_wave_blocks"   seed blocks per wave
 *   "lanes"            waves in flight, 1 .. the number of lanes the engine was created with
 *                      (environment FCX_LANES at creation, default 3); 0 = all.  With 1 the per-kernel
 *                      timings of fcx_last_stats are not inflated by overlap
 *   "dp_variant"       3 = k_dp3 (default: diagonals pinned to lanes, V in registers), 1 = k_dp (round-1
 *                      kernel, V ring in shared memory), 2 = k_dp with the spans staged by TMA
 *   "profile"          0/1
 *   "debug_split_above", "debug_tiny_capacity"   TEST HOOKS: simulate an out-of-memory wave split / start
 *                      every wave with arenas that are too small (exercises the retry paths)
 * Environment read at fcx_create: FCX_LANES, FCX_ARENA_GB, FCX_WAVE_BLOCKS, FCX_WAVE_PAIRS, FCX_DP_VARIANT,
 * FCX_PROFILE, FCX_TRACE_WAVES (per-wave host timeline on stderr). */
int fcx_set_option(fcx_ctx *, const char *name, double value);

/* Batched banded alignment of sequences of the uploaded pool, distance only: the per-pair call of
 * falcon_kit/mains/graph_to_contig.py:50-103 -- DWA.align(q[s1:e1], e1-s1, t[s2:e2], e2-s2, 1500, 1),
 * of which the caller keeps aln_str_size and dist -- for n pairs at once, one warp per pair.
 * ranges: 4 ints (s1, e1, s2, e2) per pair, or NULL for whole sequences.  Results are those of the
 * reference's align() (src/c/DW_banded.c:115-330): aln_str_size 0 means "not aligned". */
typedef struct { int32_t aln_str_size, dist, aln_q_e, aln_t_e; } fcx_align_result;
int fcx_align_pairs(fcx_ctx *, uint32_t n_pairs, const uint32_t *q_ids, const uint32_t *t_ids,
                    const int32_t *ranges, int band_tolerance, fcx_align_result *out);

/* --trim (falcon_kit/mains/consensus.py:123-158, get_consensus_with_trim) for a batch of seed
 * blocks.  For every non-seed read the k-mer chaining of get_alignment (consensus.py:48-99:
 * mask_k_mer(.., 16), find_kmer_pos_for_seq, find_best_aln_range2(K, 400, 25); src/c/kmer_lookup.c:
 * 195-204, 207-286, 429-585) runs on the device; reads are kept / cut by the reference's rules
 * (aln_score > 1000, span > 500, edge_tolerance, trim_size off both ends), ordered longest
 * alignment first, capped like get_longest_reads(.., sort=False); the trimmed reads are cut out of
 * the packed pool on the device and APPENDED to it as new reads.  Returns new block lists in the
 * shape fcx_consensus_blocks accepts (owned by the engine, valid until the next fcx_trim_blocks)
 * and the new pool size.  fcx_pool_truncate(n) drops everything after the first n reads again. */
int fcx_trim_blocks(fcx_ctx *, uint32_t n_blocks, const uint32_t *block_off, const uint32_t *read_ids,
                    int edge_tolerance, int trim_size, unsigned max_n_read, unsigned max_cov_aln,
                    const uint32_t **out_block_off, const uint32_t **out_read_ids,
                    uint32_t *out_n_reads);
int fcx_pool_truncate(fcx_ctx *, uint32_t n_reads);

/* ---- several GPUs in one process -------------------------------------------
 * A fcx_multi owns one engine per listed device.  fcx_multi_pool_upload gives EVERY device the whole
 * 2-bit read store: device d packs 1/N of the reads from host memory, the other devices receive that
 * part by a peer copy (NVLink).  fcx_multi_consensus_blocks cuts the seed blocks into N contiguous
 * cost-balanced slices, runs slice d on device d and returns the results merged in seed order -- the
 * ordering contract of exe_pool.imap (falcon_kit/mains/consensus.py:274).  Same argument meaning,
 * ownership and error behaviour as the single-device calls. */
typedef struct fcx_multi fcx_multi;
int fcx_multi_create(const int *devices, int n_devices, fcx_multi **out);
void fcx_multi_destroy(fcx_multi *);
const char *fcx_multi_last_error(const fcx_multi *);
int fcx_multi_device_count(const fcx_multi *);
int fcx_multi_set_option(fcx_multi *, const char *name, double value);
int fcx_multi_pool_upload(fcx_multi *, const char *bases, const uint64_t *offsets, uint32_t n_reads);
uint64_t fcx_multi_peer_bytes(const fcx_multi *);   /* bytes moved GPU-to-GPU by the last pool upload */
int fcx_multi_consensus_blocks(fcx_multi *, uint32_t n_blocks, const uint32_t *block_off,
                               const uint32_t *read_ids, unsigned min_cov, unsigned K, double min_idt,
                               const char **out_bases, const uint64_t **out_off);
int fcx_multi_last_pair_info(fcx_multi *, fcx_pair_info *out, uint64_t max_pairs, uint64_t *n_pairs);
int fcx_multi_last_stats(fcx_multi *, double *times_ms, uint64_t *counters);

/* INTERNAL hooks (not part of the supported surface): used by the legacy symbols inside the library
 * (fcx_legacy.cu) and by tools/profile_run.py. */
int fcx_internal_want_eqv(fcx_ctx *, int on);
int fcx_internal_last_eqv(fcx_ctx *, const int32_t **eqv, uint64_t *n);
int fcx_internal_align(fcx_ctx *, const char *q, int q_len, const char *t, int t_len, int band_tolerance,
                       int get_aln_str, alignment *out);
int fcx_internal_profile(fcx_ctx *, double *out8);

/* CUDA-event stopwatch on the engine's stream: start records an event, stop records a second one,
 * waits for it and returns the elapsed device time in milliseconds. */
int fcx_timer_start(fcx_ctx *);
int fcx_timer_stop(fcx_ctx *, double *ms);

/* Library identification: returns "falcon_b200 0.3 sm_100a". */
const char *fcx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FALCON_B200_H */
