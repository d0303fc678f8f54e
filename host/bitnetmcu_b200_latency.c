/* bitnetmcu_b200_latency.c -- single-image latency of an Inference() export (plain C, gcc).
 *
 *     latency <dll> <digits.bin> <calls>
 *
 * dlopen()s <dll> -- the gcc-built shim of this library (host/bitnetmcu_b200_dll.c) or the reference's own
 * Bitnet_inf.dll -- and calls `uint32_t Inference(int8_t *input)` (/root/reference/BitNetMCU_MNIST_dll.c:24-26) the way
 * the reference's caller does: one 16x16 image per call (/root/reference/test_inference.py:146-150), <calls> times over the
 * 256-byte images of <digits.bin>.  Prints one JSON line: microseconds per call (after 100 warm-up calls) and the labels
 * of the first ten images. */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

typedef uint32_t (*inference_fn)(int8_t *);

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s <dll> <digits.bin> <calls>\n", argv[0]); return 2; }
    void *h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
    inference_fn fn = (inference_fn)dlsym(h, "Inference");
    if (!fn) { fprintf(stderr, "no Inference() in %s\n", argv[1]); return 1; }
    static int8_t imgs[64][256];
    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 1; }
    size_t n_img = fread(imgs, 256, 64, f);
    fclose(f);
    if (n_img == 0) return 1;
    long calls = atol(argv[3]);
    uint32_t labels[10] = {0};
    for (size_t i = 0; i < 10 && i < n_img; i++) labels[i] = fn(imgs[i]);
    for (int i = 0; i < 100; i++) fn(imgs[i % n_img]);
    struct timespec t0, t1;
    uint32_t sink = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (long i = 0; i < calls; i++) sink += fn(imgs[i % n_img]);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double us = ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 1e3 / (double)calls;
    printf("{\"us_per_call\": %.3f, \"calls\": %ld, \"labels\": \"", us, calls);
    for (size_t i = 0; i < 10 && i < n_img; i++) printf("%s%u", i ? "," : "", labels[i]);
    printf("\", \"sink\": %u}\n", sink);
    return 0;
}
