/*
 * fc_oracle.h -- CPU restatement of FALCON's pre-assembly consensus arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the
 * product (falcon_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / the CPU baseline.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_vs_ref.py) against the unmodified
 * reference sources compiled into oracle/_ref/falcon.so, on the reference's own test_data reads
 * (t1.fa / t2.fa -> tests/golden/) and on seeded synthetic blocks, stage by stage.
 *
 * Reference: /root/reference/src/c/{kmer_lookup.c,DW_banded.c,falcon.c}.
 */
#ifndef FC_ORACLE_H
#define FC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* per (seed, read j) record of what the reference's generate_consensus loop did
 * (falcon.c:597-647) */
typedef struct {
    int n_match;       /* k-mer hits (kmer_lookup.c:207-286) */
    int s1, e1, s2, e2;/* chosen ranges (kmer_lookup.c:294-427) */
    long score;
    int passed_filter; /* span filters falcon.c:613-619 */
    int aligned;       /* align() reached an end (aln_str_size > 0) */
    int dist;          /* number of indel steps D */
    int aln_size;      /* alignment columns A */
    int q_e, t_e;      /* aln_q_e, aln_t_e */
    int accepted;      /* falcon.c:629 */
    int n_tags;        /* tags kept by get_align_tags (falcon.c:106-162) */
    long trace_cells;  /* E: number of (d,k) cells evaluated (DW_banded.c:188-225) */
} orc_pair_info;

/* k-mer range for one (read, seed) pair: find_kmer_pos_for_seq + find_best_aln_range(K,6K,5) */
void orc_kmer_range(const char *read, int rlen, const char *seed, int slen, unsigned K,
                    orc_pair_info *out);

/* banded O(ND) alignment, restating DW_banded.c:115-330.  q_aln / t_aln (caller-allocated,
 * q_len+t_len+1 bytes each, may be NULL) receive the gapped strings.  Returns aln_str_size. */
int orc_align(const char *q, int q_len, const char *t, int t_len, int band_tolerance,
              int *dist, int *q_e, int *t_e, char *q_aln, char *t_aln, long *cells);

/* whole seed block, restating falcon.c:562-666.  Returns a malloc'd NUL-terminated consensus
 * (caller frees with orc_free).  info: n_seq entries (entry 0 unused) or NULL. eqv_out: optional
 * malloc'd int array of strlen(result) entries. */
char *orc_generate_consensus(const char **seqs, unsigned n_seq, unsigned min_cov, unsigned K,
                             double min_idt, orc_pair_info *info, int **eqv_out);
void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
