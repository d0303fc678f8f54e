/*
 * ref_driver.c -- batch driver around the UNMODIFIED reference kernels (test infrastructure).
 *
 * This file is ours; it is linked against /root/reference/BitNetMCU_inference.c compiled where
 * it lies (oracle/Makefile, target `ref`), and only strings the reference's own exported
 * functions together the way BitMnistInference does (BitNetMCU_MNIST_dll.c:48-121), driven by a
 * runtime layer table instead of the compile-time BitNetMCU_model.h macros, so one build serves
 * every model header.  It additionally hands back the int32 logits the reference computes but
 * never returns.  Used (a) to pin oracle/bitnet_oracle.c, (b) to generate
 * tests/golden/, (c) as bench.py's cpu_baseline / --impl reference arm ("kind": "reference").
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#include "BitNetMCU_inference.h" /* from -I/root/reference: the reference's own prototypes */

typedef struct ref_layer {
    uint32_t kind; /* 0 FC, 1 conv33, 2 maxpool22 */
    int32_t bitperweight;
    uint32_t n_in;
    uint32_t n_out;
    const void *weights;
} ref_layer;

#define REF_MAX_ACT 4096

/* FC chain, dll.c:95-121 (generalised from 3/4 layers to n). */
static uint32_t ref_fc_tail(const ref_layer *fc, uint32_t n_fc, int8_t *layer_in, int32_t *layer_out, int32_t *logits)
{
    uint32_t label = 255;
    for (uint32_t l = 0; l < n_fc; l++) {
        processfclayer(layer_in, (const uint32_t *)fc[l].weights, fc[l].bitperweight, fc[l].n_in, fc[l].n_out, layer_out);
        if (l + 1 == n_fc) memcpy(logits, layer_out, sizeof(int32_t) * fc[l].n_out);
        label = ReLUNorm(layer_out, layer_in, fc[l].n_out);
    }
    return label;
}

static uint32_t ref_one(int model_class, const ref_layer *layers, uint32_t n_layers, const int8_t *img,
                        uint32_t img_bytes, int32_t *logits)
{
    int32_t layer_out[REF_MAX_ACT];
    /* int32-aligned: the CNN path aliases it as the int32 feature vector (dll.c:50,63) */
    int32_t layer_in32[REF_MAX_ACT];
    int8_t *layer_in = (int8_t *)layer_in32;
    const ref_layer *fc[16];
    ref_layer fcl[16];
    uint32_t n_fc = 0;
    const ref_layer *conv[3];
    uint32_t n_conv = 0;
    for (uint32_t l = 0; l < n_layers; l++) {
        if (layers[l].kind == 0 && n_fc < 16) fc[n_fc++] = &layers[l];
        if (layers[l].kind == 1 && n_conv < 3) conv[n_conv++] = &layers[l];
    }
    for (uint32_t l = 0; l < n_fc; l++) fcl[l] = *fc[l];
    memset(layer_in32, 0, sizeof(layer_in32));

    if (model_class == 1) {
        /* dll.c:64-80; scratch is 32x32 here because MAX_N_ACTIVATIONS < 256 in the 48/32/16-channel
         * headers would overflow the reference's own stack buffer. */
        int32_t tmpbuf[32 * 32];
        int32_t *outputptr = layer_in32;
        uint32_t xy = conv[0]->n_in, channels = conv[0]->n_out;
        for (uint32_t ch = 0; ch < channels; ch++) {
            for (uint32_t i = 0; i < xy * xy; i++) tmpbuf[i] = img[i];
            processconv33ReLU(tmpbuf, (const int8_t *)conv[0]->weights + 9 * ch, xy, 4, tmpbuf);
            processconv33ReLU(tmpbuf, (const int8_t *)conv[1]->weights + 9 * ch, xy - 2, 4, tmpbuf);
            processmaxpool22(tmpbuf, xy - 4, tmpbuf);
            processconv33ReLU(tmpbuf, (const int8_t *)conv[2]->weights + 9 * ch, (xy - 4) / 2, 4, tmpbuf);
            outputptr = processmaxpool22(tmpbuf, (xy - 4) / 2 - 2, outputptr);
        }
        ReLUNorm(layer_in32, layer_in, (uint32_t)(outputptr - layer_in32));
    } else {
        memcpy(layer_in, img, img_bytes);
    }
    return ref_fc_tail(fcl, n_fc, layer_in, layer_out, logits);
}

typedef struct {
    int model_class;
    const ref_layer *layers;
    uint32_t n_layers;
    const int8_t *images;
    uint32_t img_bytes, n_classes;
    int32_t *logits;
    uint32_t *labels;
    size_t begin, end;
} ref_job;

static void *ref_worker(void *arg)
{
    ref_job *j = (ref_job *)arg;
    for (size_t b = j->begin; b < j->end; b++) {
        uint32_t lab = ref_one(j->model_class, j->layers, j->n_layers, j->images + b * j->img_bytes, j->img_bytes,
                               j->logits + b * j->n_classes);
        if (j->labels) j->labels[b] = lab;
    }
    return NULL;
}

int ref_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* images fanned over pthreads: the reference kernels are re-entrant and stateless */
int ref_infer_batch(int model_class, const ref_layer *layers, uint32_t n_layers, const int8_t *images, size_t n,
                    uint32_t img_bytes, uint32_t n_classes, int32_t *logits, uint32_t *labels, int nthreads)
{
    if (nthreads <= 0) nthreads = ref_num_threads();
    if (nthreads > 256) nthreads = 256;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t tid[256];
    ref_job job[256];
    for (int t = 0; t < nthreads; t++) {
        job[t] = (ref_job){model_class, layers, n_layers, images, img_bytes, n_classes, logits, labels,
                           n * t / nthreads, n * (t + 1) / nthreads};
        if (t + 1 < nthreads) pthread_create(&tid[t], NULL, ref_worker, &job[t]);
    }
    ref_worker(&job[nthreads - 1]);
    for (int t = 0; t + 1 < nthreads; t++) pthread_join(tid[t], NULL);
    return 0;
}
