/*
 * The C ABI from plain C (no Python, no torch): load a genes table and its features table the way `gecco predict`
 * does (gecco/cli/commands/predict.py:62-110), pack the batch a device call takes, write the genes table back.
 * With a B200 present the same handle-free calls are followed by gcrf_model_create / gcrf_marginals_windowed
 * (GCRF_FLAG_ACCESSIONS) — see INTEGRATION.md; this file only needs the host part, so it also runs without a GPU.
 *
 *   gcc -std=c99 -Iinclude examples/tables_roundtrip.c -Lgecco_b200 -lgecco_crf_b200 -Wl,-rpath,$PWD/gecco_b200 -lm
 *   ./a.out x.genes.tsv x.features.tsv out.genes.tsv
 */
#include <math.h>
#include <stdio.h>

#include "gecco_crf_b200.h"

int main(int argc, char **argv) {
    if (argc != 4) {
        fprintf(stderr, "usage: %s genes.tsv features.tsv out.genes.tsv\n", argv[0]);
        return 2;
    }
    gcrf_table *table = NULL;
    const char *features[1] = {argv[2]};
    if (gcrf_table_load(argv[1], features, 1, NAN /* no e-value filter */, 1e-9 /* gecco predict's p-value filter */, &table) != GCRF_OK) {
        fprintf(stderr, "load: %s\n", gcrf_table_last_error());
        return 1;
    }
    const int32_t *contig_ptr = NULL, *row_ptr = NULL, *accession = NULL;
    int64_t rows = 0, nnz = 0;
    if (gcrf_table_pack_accessions(table, 0 /* protein */, 5 /* PF + five digits */, &contig_ptr, &row_ptr, &accession, &rows, &nnz) != GCRF_OK) {
        fprintf(stderr, "pack: %s\n", gcrf_table_last_error());
        return 1;
    }
    long long in_vocabulary_form = 0;
    for (int64_t i = 0; i < nnz; ++i) in_vocabulary_form += accession[i] >= 0;
    printf("abi=%d devices=%d contigs=%lld genes=%lld domains=%lld rows=%lld nnz=%lld pfam=%lld first_contig=%s last_row_end=%d\n",
           gcrf_version(), gcrf_device_count(), (long long)gcrf_table_contigs(table), (long long)gcrf_table_genes(table),
           (long long)gcrf_table_domains(table), (long long)rows, (long long)nnz, in_vocabulary_form,
           gcrf_table_contigs(table) ? gcrf_table_contig_id(table, 0) : "-", rows ? row_ptr[rows] : 0);
    if (gcrf_table_write_genes(table, NULL /* no probabilities yet */, argv[3]) != GCRF_OK) {
        fprintf(stderr, "write: %s\n", gcrf_table_last_error());
        return 1;
    }
    gcrf_table_destroy(table);
    return 0;
}
