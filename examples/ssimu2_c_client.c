/*
 * ssimu2_c_client.c -- a plain-C caller of libssimu2_b200.so: nothing but include/ssimu2_b200.h, no CUDA headers, no Python.
 *
 *   ssimu2_c_client <width> <height> <ref.rgb> <dis.rgb> [n_devices]
 *
 * Reads two packed sRGB8 images (width*height*3 bytes each), scores them through the host-frame entry point
 * (the analogue of Ssimulacra2::compute_from_cpu_srgb_sync, crates/ssimulacra2-cuda/src/lib.rs:232-250) and prints the
 * SSIMULACRA2 score with 17 significant digits.  With n_devices > 1 the pair is scored once per GPU through ssimu2_shard_*
 * (one handle + host thread per device inside the library) and every score is printed.
 *
 * Build:  gcc -O2 -Iinclude examples/ssimu2_c_client.c -Lturbo_metrics_b200 -lssimu2_b200 -Wl,-rpath,$PWD/turbo_metrics_b200 -o ssimu2_c_client
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ssimu2_b200.h"

static unsigned char *read_file(const char *path, size_t n)
{
    FILE *f = fopen(path, "rb");
    unsigned char *p = (unsigned char *)malloc(n);
    if (!f || !p || fread(p, 1, n, f) != n) {
        fprintf(stderr, "cannot read %zu bytes from %s\n", n, path);
        exit(2);
    }
    fclose(f);
    return p;
}

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != SSIMU2_OK) {                                                      \
            fprintf(stderr, "%s: %s (%d)\n", #call, ssimu2_strerror(rc_), rc_);      \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s width height ref.rgb dis.rgb [n_devices]\n", argv[0]);
        return 2;
    }
    const uint32_t w = (uint32_t)atoi(argv[1]), h = (uint32_t)atoi(argv[2]);
    const int ndev = argc > 5 ? atoi(argv[5]) : 1;
    const size_t bytes = (size_t)w * h * 3;
    unsigned char *ref = read_file(argv[3], bytes), *dis = read_file(argv[4], bytes);

    ssimu2_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.width = w;
    cfg.height = h;
    cfg.format = SSIMU2_FMT_SRGB8;
    ssimu2_frame fr = {{(uint64_t)(uintptr_t)ref, 0}, w * 3, 0}, fd = {{(uint64_t)(uintptr_t)dis, 0}, w * 3, 0};

    if (ndev <= 1) {
        ssimu2_t *m = NULL;
        uint64_t ticket = 0;
        double score = 0.0, norms[108];
        CHECK(ssimu2_create(&m, &cfg));
        CHECK(ssimu2_submit_host(m, &fr, &fd, bytes, &ticket));
        CHECK(ssimu2_get_score(m, ticket, &score));
        CHECK(ssimu2_get_norms(m, ticket, norms));
        printf("%.17g\n", score);
        CHECK(ssimu2_destroy(m));
    } else {
        ssimu2_shard_t *s = NULL;
        int32_t devices[64];
        ssimu2_frame refs[64], diss[64];
        double scores[64];
        uint64_t first = 0;
        if (ndev > 64) return 2;
        cfg.batch = 1; /* one pair per launch group, so that pair i goes to device i */
        for (int i = 0; i < ndev; i++) {
            devices[i] = i;
            refs[i] = fr;
            diss[i] = fd;
        }
        CHECK(ssimu2_shard_create(&s, &cfg, devices, (uint32_t)ndev));
        CHECK(ssimu2_shard_submit_host(s, (uint32_t)ndev, refs, diss, bytes, &first));
        CHECK(ssimu2_shard_get_scores(s, first, (uint32_t)ndev, scores));
        for (int i = 0; i < ndev; i++) printf("%.17g\n", scores[i]);
        CHECK(ssimu2_shard_destroy(s));
    }
    free(ref);
    free(dis);
    return 0;
}
