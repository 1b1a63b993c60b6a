/* A plain-C99 consumer of include/prosail_b200.h (test infrastructure): proves that the drop-in boundary is a C ABI -- no C++,
 * no Python, no torch on the caller's side.  Reads the coefficient tables and the per-band interface constants from a raw binary
 * file written by tests/test_c_abi.py, evaluates pyrism.PROSPECT (models.py:900-921) for a few parameter sets given as HOST
 * pointers, and writes ks, kt ([n][2101] doubles each) to the output file.
 *   usage: prospect_demo <tables.bin> <out.bin>      exit 0 and "NOGPU" on stdout when there is no B200 (no CPU fallback).
 * tables.bin: float32 Kab Kxc Kbr Kw Km [2101 each], float64 talf t12 t21 [2101 each], int32 n, float64 N Cab Cxc Cbr Cw Cm [n each] */
#include <stdio.h>
#include <stdlib.h>

#include "prosail_b200.h"

#define NB PROSAIL_NBANDS

static int fail(const char* what, int rc) {
    fprintf(stderr, "%s failed: rc=%d (%s)\n", what, rc, prosail_last_error());
    return 1;
}

int main(int argc, char** argv) {
    prosail_handle_t* h = NULL;
    int rc = prosail_create(0, &h);
    if (rc == PROSAIL_E_NOGPU) {
        printf("NOGPU %s\n", prosail_version());
        return 0;
    }
    if (rc) return fail("prosail_create", rc);
    if (argc < 3) { fprintf(stderr, "usage: prospect_demo tables.bin out.bin\n"); return 2; }

    static float K[5][NB];
    static double T[3][NB];
    int n = 0;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    if (fread(K, sizeof(float), 5 * NB, f) != 5 * NB || fread(T, sizeof(double), 3 * NB, f) != 3 * NB || fread(&n, sizeof(int), 1, f) != 1 || n < 1) {
        fprintf(stderr, "short tables file\n");
        return 2;
    }
    double* P = (double*)malloc(sizeof(double) * 6 * (size_t)n);
    double* ks = (double*)malloc(sizeof(double) * NB * (size_t)n);
    double* kt = (double*)malloc(sizeof(double) * NB * (size_t)n);
    if (!P || !ks || !kt || fread(P, sizeof(double), 6 * (size_t)n, f) != 6 * (size_t)n) { fprintf(stderr, "short parameter block\n"); return 2; }
    fclose(f);

    if ((rc = prosail_set_leaf_tables(h, PROSAIL_VERSION_5, K[0], K[1], K[2], K[3], K[4], NULL, T[0], T[1], T[2]))) return fail("prosail_set_leaf_tables", rc);

    prosail_leaf_t leaf = {P, P + n, P + 2 * n, P + 3 * n, P + 4 * n, P + 5 * n, NULL};
    prosail_out_t out = {{0}, 0, 0, NULL, NULL, 0};
    out.plane[PROSAIL_O_LEAF_R] = ks;
    out.plane[PROSAIL_O_LEAF_T] = kt;
    out.pitch = NB;
    if ((rc = prosail_prospect_f64(h, PROSAIL_VERSION_5, n, &leaf, &out, NULL, /*on_host=*/1, NULL))) return fail("prosail_prospect_f64", rc);

    /* argument errors come back as codes, never as exceptions or aborts */
    if (prosail_prospect_f64(h, 7, n, &leaf, &out, NULL, 1, NULL) != PROSAIL_E_ARG) { fprintf(stderr, "bad version was accepted\n"); return 1; }

    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 2; }
    fwrite(ks, sizeof(double), NB * (size_t)n, f);
    fwrite(kt, sizeof(double), NB * (size_t)n, f);
    fclose(f);
    printf("OK %d sets, ks[0][0]=%.17g\n", n, ks[0]);
    prosail_destroy(h);
    free(P); free(ks); free(kt);
    return 0;
}
