/*
 * The device half of the C ABI from plain C99 (no Python, no torch, no CUDA headers): a three-attribute model, two
 * contigs given as host arrays, per-gene cluster probabilities back — the call a GECCO maintainer would bind in place
 * of the per-window tagger call at gecco/crf/__init__.py:253.  The same batch then goes
 * through the reference-order f64 arithmetic (GCRF_FLAG_F64) and through the compact wire format (gcrf_wire_*).
 * Without a B200 gcrf_model_create fails with GCRF_ENODEVICE and the program says so (exit code 3): there is no CPU path.
 *
 *   gcc -std=c99 -Iinclude examples/marginals_device.c -Lgecco_b200 -lgecco_crf_b200 -Wl,-rpath,$PWD/gecco_b200 -lm
 */
#include <math.h>
#include <stdio.h>

#include "gecco_crf_b200.h"

int main(void) {
    /* state weights [A][L] (label 0 = '0', label 1 = '1'), transition weights [L][L] from -> to */
    const double state_w[3][2] = {{0.5, -0.25}, {-1.0, 2.0}, {0.0, 0.75}};
    const double trans_w[2][2] = {{2.5, -2.5}, {-2.5, 2.5}};
    /* two contigs: 3 genes (shorter than the window of 5: padded) and 6 genes; gene 4 has no domains, one id is unknown */
    const int32_t contig_ptr[3] = {0, 3, 9};
    const int32_t gene_ptr[10] = {0, 1, 3, 3, 4, 4, 6, 7, 9, 10};
    const int32_t attr_idx[10] = {1, 0, 2, 1, 2, 1, -1, 0, 1, 2};
    double p32[9], p64[9], pw[9];
    gcrf_model *model = NULL;
    int rc = gcrf_model_create(&state_w[0][0], 3, 2, &trans_w[0][0], 1 /* report the marginal of label '1' */, 0, &model);
    if (rc == GCRF_ENODEVICE) {
        printf("no usable B200: %s\n", gcrf_last_error());
        return 3;
    }
    if (rc != GCRF_OK) {
        fprintf(stderr, "gcrf_model_create: %s\n", gcrf_last_error());
        return 1;
    }
    rc = gcrf_marginals_windowed(model, contig_ptr, gene_ptr, attr_idx, 2, 9, 10, 5, 1, 1, p32, 0);
    if (rc == GCRF_OK) rc = gcrf_marginals_windowed(model, contig_ptr, gene_ptr, attr_idx, 2, 9, 10, 5, 1, 1, p64, GCRF_FLAG_F64);
    gcrf_wire *wire = NULL;
    if (rc == GCRF_OK && gcrf_wire_encode(contig_ptr, gene_ptr, attr_idx, 2, 9, 10, 3, 0, &wire) != GCRF_OK) {
        fprintf(stderr, "gcrf_wire_encode: %s\n", gcrf_wire_last_error());
        return 1;
    }
    if (rc == GCRF_OK) rc = gcrf_marginals_windowed_wire(model, wire, 5, 1, 1, pw, 0);
    if (rc != GCRF_OK) {
        fprintf(stderr, "marginals: %s\n", gcrf_last_error());
        return 1;
    }
    double worst = 0.0;
    int same = 1;
    for (int g = 0; g < 9; ++g) {
        printf("gene %d  p=%.9f  f64=%.17g\n", g, p32[g], p64[g]);
        if (fabs(p32[g] - p64[g]) > worst) worst = fabs(p32[g] - p64[g]);
        same = same && p32[g] == pw[g];
    }
    printf("max_abs_diff_fp32_vs_f64=%.3e wire_equals_csr=%d wire_bytes=%lld launches=%lld\n", worst, same,
           (long long)gcrf_wire_bytes(wire), (long long)gcrf_model_launch_count(model));
    gcrf_wire_destroy(wire);
    gcrf_model_destroy(model);
    return worst <= 1e-5 && same ? 0 : 1;
}
