/* Plain-C caller of the legacy ABI: what a C program linked against the reference's falcon.so
 * would do (src/c/falcon.c:562-566, :776-780; the dead main() at falcon.c:782-840 documents the same
 * calling sequence).  Reads one sequence per line from argv[1] (first line = seed), calls
 * generate_consensus() and prints the consensus.  Built and run by tests/test_c_abi.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "falcon_b200.h"

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s seqs.txt min_cov min_idt\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "r");
    if (!f) { perror("open"); return 2; }
    char **seqs = NULL; unsigned n = 0, cap = 0;
    char *line = NULL; size_t lcap = 0; ssize_t len;
    while ((len = getline(&line, &lcap, f)) > 0) {
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len == 0) continue;
        if (n == cap) { cap = cap ? cap * 2 : 64; seqs = realloc(seqs, cap * sizeof(char *)); }
        seqs[n++] = strdup(line);
    }
    fclose(f);
    consensus_data *cd = generate_consensus(seqs, n, (unsigned)atoi(argv[2]), 8, atof(argv[3]));
    printf("%s\n", cd->sequence);
    long eqv_sum = 0;
    for (size_t i = 0; i < strlen(cd->sequence); i++) eqv_sum += cd->eqv[i];
    fprintf(stderr, "len=%zu eqv_sum=%ld\n", strlen(cd->sequence), eqv_sum);
    free_consensus_data(cd);
    for (unsigned i = 0; i < n; i++) free(seqs[i]);
    free(seqs); free(line);
    return 0;
}
