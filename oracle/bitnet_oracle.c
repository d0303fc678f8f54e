/*
 * bitnet_oracle.c -- CPU ORACLE (test infrastructure; see bitnet_oracle.h for the rules).
 *
 * Restates /root/reference/BitNetMCU_inference.c and the BitMnistInference chains of
 * /root/reference/BitNetMCU_MNIST_dll.c.  Every function cites the reference lines it follows.
 * Written table-driven: a packed weight word is first decoded to signed integers, then a plain
 * dot product is taken.  That is arithmetically identical to the reference's bit-serial loops
 * (all-integer, int32 accumulate, no overflow at the sizes of SURVEY.md 8a).
 */
#include "bitnet_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * Weight decode.  One table per encoding id, indexed by the code taken from the MSB side of the
 * word (the reference tests bit 31 and shifts left: inference.c:96-201).
 * ---------------------------------------------------------------------------------------- */

/* id 4 "4bitsym": bit3 = sign (set = negative), low 3 bits m -> +-(2m+1)   (inference.c:156-168) */
static const int16_t k_lut_4bitsym[16] = {1, 3, 5, 7, 9, 11, 13, 15, -1, -3, -5, -7, -9, -11, -13, -15};
/* id 12 "4bit": two's complement nibble                                    (inference.c:169-178) */
static const int16_t k_lut_4bit[16] = {0, 1, 2, 3, 4, 5, 6, 7, -8, -7, -6, -5, -4, -3, -2, -1};
/* id 20 "FP130": bit3 = sign, low 3 bits e -> +-(1<<e)                     (inference.c:190-201) */
static const int16_t k_lut_fp130[16] = {1, 2, 4, 8, 16, 32, 64, 128, -1, -2, -4, -8, -16, -32, -64, -128};
/* id 2 "2bitsym": bit1 = sign, bit0 = magnitude bit -> +-1, +-3            (inference.c:105-115) */
static const int16_t k_lut_2bitsym[4] = {1, 3, -1, -3};
/* id 36 "NF4": NOT decoded by the reference (inference.c:202 -> zeros).  Extension LUT =
 * round(127*level) of the level table in exportquant.py:117-118.  parity unpinned. */
static const int16_t k_lut_nf4_ext[16] = {-127, -88, -67, -50, -36, -23, -12, 0, 10, 20, 31, 43, 56, 71, 92, 127};

/* Decode one row of n_input weights starting at `row_words` into dense[0..n_input). */
static int decode_row(const uint32_t *row_words, const uint16_t *row_words16, int32_t enc, uint32_t n_input,
                      int16_t *dense, int nf4_extension)
{
    uint32_t k = 0;
    switch (enc) {
    case 1: /* Binary: bit set = +1, clear = -1 (inference.c:96-104) */
        for (; k < n_input; k += 32) {
            uint32_t w = *row_words++;
            for (uint32_t j = 0; j < 32 && k + j < n_input; j++) dense[k + j] = ((w >> (31 - j)) & 1u) ? 1 : -1;
        }
        return 0;
    case 2:
        for (; k < n_input; k += 16) {
            uint32_t w = *row_words++;
            for (uint32_t j = 0; j < 16 && k + j < n_input; j++) dense[k + j] = k_lut_2bitsym[(w >> (30 - 2 * j)) & 3u];
        }
        return 0;
    case 4:
    case 12:
    case 20:
    case 36: {
        const int16_t *lut = enc == 4 ? k_lut_4bitsym : enc == 12 ? k_lut_4bit : enc == 20 ? k_lut_fp130 : k_lut_nf4_ext;
        if (enc == 36 && !nf4_extension) break;
        for (; k < n_input; k += 8) {
            uint32_t w = *row_words++;
            for (uint32_t j = 0; j < 8 && k + j < n_input; j++) dense[k + j] = lut[(w >> (28 - 4 * j)) & 15u];
        }
        return 0;
    }
    case 16: /* 8-bit two's complement (inference.c:179-188) */
        for (; k < n_input; k += 4) {
            uint32_t w = *row_words++;
            for (uint32_t j = 0; j < 4 && k + j < n_input; j++) dense[k + j] = (int16_t)(int8_t)((w >> (24 - 8 * j)) & 255u);
        }
        return 0;
    case 64: /* Ternary: 10 trits per uint16, extracted by repeated *3 (inference.c:116-136):
                after c*=3, bit17 set -> weight 0, else bit16 set -> -1, clear -> +1; c&=0xFFFF */
        for (; k < n_input; k += 10) {
            uint32_t c = *row_words16++;
            for (uint32_t j = 0; j < 10; j++) {
                c *= 3u;
                int16_t w = (c & 0x20000u) ? 0 : ((c & 0x10000u) ? -1 : 1);
                if (k + j < n_input) dense[k + j] = w;
                c &= 0xFFFFu;
            }
        }
        return 0;
    default:
        break;
    }
    /* unsupported id: the reference leaves sum = 0 (inference.c:202) */
    for (k = 0; k < n_input; k++) dense[k] = 0;
    return -1;
}

/* words (uint32) per row for the 32-bit packed encodings; Ternary uses n_input/10 uint16 (inference.c:118) */
static uint32_t weights_per_word(int32_t enc)
{
    switch (enc) {
    case 1: return 32;
    case 2: return 16;
    case 4: case 12: case 20: case 36: return 8;
    case 16: return 4;
    default: return 0;
    }
}

int orc_decode_fc(const uint32_t *weights, int32_t enc, uint32_t n_input, uint32_t n_output, int16_t *dense,
                  int nf4_extension)
{
    int rc = 0;
    uint32_t wpw = weights_per_word(enc);
    for (uint32_t i = 0; i < n_output; i++) {
        const uint32_t *row32 = wpw ? weights + (size_t)i * ((n_input + wpw - 1) / wpw) : weights;
        const uint16_t *row16 = (const uint16_t *)weights + (size_t)i * (n_input / 10);
        rc |= decode_row(row32, row16, enc, n_input, dense + (size_t)i * n_input, nf4_extension);
    }
    return rc;
}

/* processfclayer, inference.c:88-208: output[i] = sum_k decode(W[i][k]) * activations[k] */
void orc_processfclayer(const int8_t *activations, const uint32_t *weights, int32_t enc, uint32_t n_input,
                        uint32_t n_output, int32_t *output)
{
    int16_t *row = (int16_t *)malloc(sizeof(int16_t) * (n_input ? n_input : 1));
    uint32_t wpw = weights_per_word(enc);
    for (uint32_t i = 0; i < n_output; i++) {
        const uint32_t *row32 = wpw ? weights + (size_t)i * ((n_input + wpw - 1) / wpw) : weights;
        const uint16_t *row16 = (const uint16_t *)weights + (size_t)i * (n_input / 10);
        decode_row(row32, row16, enc, n_input, row, 0);
        int32_t sum = 0;
        for (uint32_t k = 0; k < n_input; k++)
            if (row[k]) sum += (int32_t)row[k] * (int32_t)activations[k]; /* zero weights never touch the activation
                                                                            (ternary padding, inference.c:128) */
        output[i] = sum;
    }
    free(row);
}

/* ReLUNorm, inference.c:23-72.
 *  (1) argmax = first index of the maximum, strict '>', start value -INT32_MAX, position 255 (32-37)
 *  (2) shift = bit length of (max >> 7) (41-47); rounding = (1<<shift)>>1 (51)
 *  (3) x<0 -> 0 else min(127,(x+rounding)>>shift) (54-67)
 * For max<0 the C computes shift=32 and an undefined 1<<32; every element is then negative so all
 * outputs are 0 -- restated here as that.  */
uint32_t orc_ReLUNorm(const int32_t *input, int8_t *output, uint32_t n_input)
{
    int32_t max_val = -INT32_MAX;
    uint32_t max_pos = 255;
    for (uint32_t i = 0; i < n_input; i++)
        if (input[i] > max_val) { max_val = input[i]; max_pos = i; }

    uint32_t shift = 0;
    if (max_val > 0) {
        uint32_t scale = (uint32_t)(max_val >> 7);
        while (scale) { shift++; scale >>= 1; }
    }
    int32_t rounding = (int32_t)((1u << shift) >> 1);
    for (uint32_t i = 0; i < n_input; i++) {
        int32_t x = input[i];
        if (x < 0) { output[i] = 0; continue; }
        int32_t t = (x + rounding) >> shift;
        output[i] = (int8_t)(t > 127 ? 127 : t);
    }
    return max_pos;
}

/* processconv33ReLU, inference.c:238-277: valid 3x3 cross-correlation, one channel, stride 1;
 * s<0 -> 0 else s>>n_shift (no rounding, no clip: 261-272).  Dense (xy-2)^2 output, returns end.
 * `output` may alias `activations` (the write index never overtakes the reads). */
int32_t *orc_processconv33ReLU(const int32_t *act, const int8_t *w, uint32_t xy, uint32_t n_shift, int32_t *output)
{
    int32_t k[9];
    for (int i = 0; i < 9; i++) k[i] = w[i];
    for (uint32_t y = 0; y + 2 < xy; y++)
        for (uint32_t x = 0; x + 2 < xy; x++) {
            const int32_t *p = act + y * xy + x;
            int32_t s = 0;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) s += k[3 * r + c] * p[r * xy + c];
            *output++ = s < 0 ? 0 : (s >> n_shift);
        }
    return output;
}

/* processmaxpool22, inference.c:300-322 */
int32_t *orc_processmaxpool22(const int32_t *act, uint32_t xy, int32_t *output)
{
    uint32_t o = xy / 2;
    for (uint32_t y = 0; y < o; y++)
        for (uint32_t x = 0; x < o; x++) {
            const int32_t *p = act + (2 * y) * xy + 2 * x;
            int32_t m = p[0];
            if (p[xy] > m) m = p[xy];
            if (p[1] > m) m = p[1];
            if (p[xy + 1] > m) m = p[xy + 1];
            *output++ = m;
        }
    return output;
}

/* ------------------------------------------------------------------------------------------
 * Whole-model chains (BitNetMCU_MNIST_dll.c:48-121) over a batch, images fanned over threads.
 * ---------------------------------------------------------------------------------------- */
#define ORC_MAX_FC 16
#define ORC_MAX_ACT 4096

typedef struct {
    uint32_t n_fc;
    uint32_t n_in[ORC_MAX_FC], n_out[ORC_MAX_FC];
    int16_t *dense[ORC_MAX_FC];
    /* CNN front-end (dll.c:64-80): conv(L2) conv(L4) pool(L6) conv(L7) pool(L9) per channel */
    int has_cnn;
    uint32_t channels, xy0;
    const int8_t *cw[3];
} orc_plan;

static void fc_dense(const int16_t *dense, uint32_t n_in, uint32_t n_out, const int8_t *act, int32_t *out)
{
    for (uint32_t i = 0; i < n_out; i++) {
        const int16_t *r = dense + (size_t)i * n_in;
        int32_t s = 0;
        for (uint32_t k = 0; k < n_in; k++) s += (int32_t)r[k] * (int32_t)act[k];
        out[i] = s;
    }
}

static uint32_t infer_one(const orc_plan *p, const int8_t *img, uint32_t img_bytes, int32_t *logits)
{
    int32_t acc[ORC_MAX_ACT];
    int8_t act[ORC_MAX_ACT];
    memset(act, 0, sizeof(act)); /* activations past a layer's real width read as 0 (ternary padding) */

    if (p->has_cnn) {
        /* dll.c:64-80.  n_shift is the literal 4 at every call site (dll.c:71-74). */
        int32_t tmp[32 * 32];
        int32_t feat[ORC_MAX_ACT];
        int32_t *fp = feat;
        uint32_t xy = p->xy0;
        for (uint32_t ch = 0; ch < p->channels; ch++) {
            for (uint32_t i = 0; i < xy * xy; i++) tmp[i] = img[i];
            orc_processconv33ReLU(tmp, p->cw[0] + 9 * ch, xy, 4, tmp);
            orc_processconv33ReLU(tmp, p->cw[1] + 9 * ch, xy - 2, 4, tmp);
            orc_processmaxpool22(tmp, xy - 4, tmp);
            orc_processconv33ReLU(tmp, p->cw[2] + 9 * ch, (xy - 4) / 2, 4, tmp);
            fp = orc_processmaxpool22(tmp, (xy - 4) / 2 - 2, fp);
        }
        orc_ReLUNorm(feat, act, (uint32_t)(fp - feat)); /* dll.c:80 */
    } else {
        memcpy(act, img, img_bytes);
    }

    uint32_t label = 255;
    for (uint32_t l = 0; l < p->n_fc; l++) {
        fc_dense(p->dense[l], p->n_in[l], p->n_out[l], act, acc);
        if (l + 1 == p->n_fc) memcpy(logits, acc, sizeof(int32_t) * p->n_out[l]);
        memset(act, 0, sizeof(act));
        label = orc_ReLUNorm(acc, act, p->n_out[l]); /* dll.c:100,106,113,116: last call's argmax is returned */
    }
    return label;
}

int orc_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const orc_plan *p;
    const int8_t *images;
    uint32_t img_bytes, n_classes;
    int32_t *logits;
    uint32_t *labels;
    size_t begin, end;
} orc_job;

static void *orc_worker(void *arg)
{
    orc_job *j = (orc_job *)arg;
    for (size_t b = j->begin; b < j->end; b++) {
        uint32_t lab = infer_one(j->p, j->images + b * j->img_bytes, j->img_bytes, j->logits + b * j->n_classes);
        if (j->labels) j->labels[b] = lab;
    }
    return NULL;
}

int orc_infer_batch(int model_class, const orc_layer *layers, uint32_t n_layers, const int8_t *images, size_t n,
                    uint32_t img_bytes, int32_t *logits, uint32_t *labels, int nthreads, int nf4_extension)
{
    orc_plan p;
    memset(&p, 0, sizeof(p));
    uint32_t n_conv = 0;
    for (uint32_t l = 0; l < n_layers; l++) {
        const orc_layer *L = &layers[l];
        if (L->kind == ORC_LAYER_FC) {
            if (p.n_fc == ORC_MAX_FC || L->n_in > ORC_MAX_ACT || L->n_out > ORC_MAX_ACT) return -1;
            p.n_in[p.n_fc] = L->n_in;
            p.n_out[p.n_fc] = L->n_out;
            p.dense[p.n_fc] = (int16_t *)calloc((size_t)L->n_in * L->n_out + 1, sizeof(int16_t));
            orc_decode_fc((const uint32_t *)L->weights, L->bitperweight, L->n_in, L->n_out, p.dense[p.n_fc], nf4_extension);
            p.n_fc++;
        } else if (L->kind == ORC_LAYER_CONV33) {
            if (n_conv == 3) return -1;
            if (n_conv == 0) { p.xy0 = L->n_in; p.channels = L->n_out; }
            p.cw[n_conv++] = (const int8_t *)L->weights;
        }
    }
    if (model_class == ORC_MODEL_CNNMNIST) {
        if (n_conv != 3 || p.xy0 > 32 || p.channels * 4 > ORC_MAX_ACT) return -1;
        p.has_cnn = 1;
    }
    if (p.n_fc == 0) return -1;
    uint32_t n_classes = p.n_out[p.n_fc - 1];

    /* images are independent and the kernels re-entrant: static slices over pthreads */
    if (nthreads <= 0) nthreads = orc_num_threads();
    if (nthreads > 256) nthreads = 256;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    {
        pthread_t tid[256];
        orc_job job[256];
        for (int t = 0; t < nthreads; t++) {
            job[t] = (orc_job){&p, images, img_bytes, n_classes, logits, labels, n * t / nthreads, n * (t + 1) / nthreads};
            if (t + 1 < nthreads) pthread_create(&tid[t], NULL, orc_worker, &job[t]);
        }
        orc_worker(&job[nthreads - 1]);
        for (int t = 0; t + 1 < nthreads; t++) pthread_join(tid[t], NULL);
    }
    for (uint32_t l = 0; l < p.n_fc; l++) free(p.dense[l]);
    return 0;
}

/* pixels = low byte of xorshift32 (s^=s<<13; s^=s>>17; s^=s<<5), one draw per pixel */
void orc_xorshift_fill(int8_t *dst, size_t n_bytes, uint32_t seed)
{
    uint32_t s = seed;
    for (size_t i = 0; i < n_bytes; i++) {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        dst[i] = (int8_t)(s & 0xFFu);
    }
}
