# ThreadSanitizer (SANITIZE=address,undefined for ASan + UBSan) over the native table path (gcrf_tables.cpp is host-only C++: built here without the CUDA part).
# Usage: bash tools/tsan_tables.sh genes.tsv features.tsv      (e.g. the tables tools/tables_time.py generates)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
cat > $TMP/stubs.c <<'STUB'
int gcrf_version(void) { return 1; }
int gcrf_device_count(void) { return 0; }
STUB
cat > $TMP/main.c <<'MAIN'
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "gecco_crf_b200.h"
int main(int argc, char **argv) {
    gcrf_table *t = NULL;
    const char *features[1] = {argc > 2 ? argv[2] : ""};
    if (gcrf_table_load(argv[1], features, argc > 2, NAN, 1e-9, &t) != GCRF_OK) { fprintf(stderr, "load: %s\n", gcrf_table_last_error()); return 1; }
    const int32_t *cp, *rp, *acc; int64_t rows, nnz;
    if (gcrf_table_pack_accessions(t, 0, 5, &cp, &rp, &acc, &rows, &nnz) != GCRF_OK) return 1;
    double *prob = malloc(sizeof(double) * (rows ? rows : 1));
    for (int64_t i = 0; i < rows; ++i) prob[i] = (double)(i % 1000) / 1000.0;
    char path[512];
    snprintf(path, sizeof path, "%s.tsan.genes.tsv", argv[1]);
    if (gcrf_table_write_genes(t, prob, path) != GCRF_OK) return 1;
    snprintf(path, sizeof path, "%s.tsan.features.tsv", argv[1]);
    if (gcrf_table_write_features(t, prob, path) != GCRF_OK) return 1;
    if (gcrf_table_pack_accessions(t, 1, 5, &cp, &rp, &acc, &rows, &nnz) != GCRF_OK) return 1;
    printf("ok contigs=%lld genes=%lld domains=%lld rows(domain mode)=%lld id0=%s\n", (long long)gcrf_table_contigs(t),
           (long long)gcrf_table_genes(t), (long long)gcrf_table_domains(t), (long long)rows, gcrf_table_contig_id(t, 0));
    gcrf_table_destroy(t);
    free(prob);
    return 0;
}
MAIN
g++ -std=c++17 -O1 -g -fsanitize=${SANITIZE:-thread} -fPIC -pthread -c $ROOT/gecco_b200/csrc/gcrf_tables.cpp -o $TMP/tables.o
gcc -std=c99 -O1 -g -fsanitize=${SANITIZE:-thread} -I$ROOT/include -c $TMP/main.c -o $TMP/main.o
gcc -c $TMP/stubs.c -o $TMP/stubs.o
g++ -fsanitize=${SANITIZE:-thread} -pthread $TMP/tables.o $TMP/main.o $TMP/stubs.o -o $TMP/tsan_tables -lm
GCRF_TABLE_THREADS=${GCRF_TABLE_THREADS:-8} $TMP/tsan_tables "$@"
