/*
 * bitnet_oracle.h -- CPU ORACLE for the BitNetMCU integer inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bitnetmcu_b200/ or include/ may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * It is a restatement (not a copy) of the reference algorithm in
 *   /root/reference/BitNetMCU_inference.c   (processfclayer 88-208, ReLUNorm 23-72,
 *                                            processconv33ReLU 238-277, processmaxpool22 300-322)
 *   /root/reference/BitNetMCU_MNIST_dll.c   (BitMnistInference FC 95-121, CNN 48-91)
 * written table-driven (decode packed word -> dense int16 row, then multiply-add)
 * instead of the reference's bit-serial loops.  Parity is PINNED: tests/test_oracle_*.py
 * check it against (a) the reference's own label KAT (BitNetMCU_MNIST_test_data.h),
 * (b) int32 logits produced by the unmodified reference compiled into oracle/_ref/
 * (fixtures under tests/golden/), (c) exhaustive per-word decode sweeps against oracle/_ref.
 */
#ifndef BITNET_ORACLE_H
#define BITNET_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* layer kinds of the runtime model descriptor (same numbering as include/bitnetmcu_b200.h) */
enum { ORC_LAYER_FC = 0, ORC_LAYER_CONV33 = 1, ORC_LAYER_MAXPOOL22 = 2 };
enum { ORC_MODEL_FCMNIST = 0, ORC_MODEL_CNNMNIST = 1 };

typedef struct orc_layer {
    uint32_t kind;          /* ORC_LAYER_*                                              */
    int32_t  bitperweight;  /* FC: encoding id 1,2,4,12,16,20,36,64 ; conv: 8          */
    uint32_t n_in;          /* FC: Lk_incoming_weights ; conv/pool: Lk_incoming_x       */
    uint32_t n_out;         /* FC: Lk_outgoing_weights ; conv: Lk_out_channels ; pool:0 */
    const void *weights;    /* FC: uint32 words (uint16 for Ternary) ; conv: int8[C*9]  */
} orc_layer;

/* --- the four reference kernels, same signatures as BitNetMCU_inference.h:15-60 --- */
uint32_t orc_ReLUNorm(const int32_t *input, int8_t *output, uint32_t n_input);
void     orc_processfclayer(const int8_t *activations, const uint32_t *weights, int32_t bits_per_weight,
                            uint32_t n_input, uint32_t n_output, int32_t *output);
int32_t *orc_processconv33ReLU(const int32_t *activations, const int8_t *weights, uint32_t xy_input,
                               uint32_t n_shift, int32_t *output);
int32_t *orc_processmaxpool22(const int32_t *activations, uint32_t xy_input, int32_t *output);

/* decode one FC layer into dense int16 [n_output][n_input] (FP130 needs +128, hence int16).
 * returns 0, or -1 for an encoding the reference does not decode (then dense is all zero,
 * exactly like inference.c:202).  nf4_extension!=0 additionally decodes id 36 with the
 * documented non-reference LUT. */
int orc_decode_fc(const uint32_t *weights, int32_t bits_per_weight, uint32_t n_input, uint32_t n_output,
                  int16_t *dense, int nf4_extension);

/* whole-model inference of n images (int8 [n][img_bytes]); logits int32 [n][n_classes];
 * labels uint32 [n] (may be NULL).  nthreads<=0: all cores.  Returns 0 on success. */
int orc_infer_batch(int model_class, const orc_layer *layers, uint32_t n_layers,
                    const int8_t *images, size_t n, uint32_t img_bytes,
                    int32_t *logits, uint32_t *labels, int nthreads, int nf4_extension);

/* the synthetic KAT stream of SURVEY.md 8(c): low byte of xorshift32, seed given, one draw per pixel */
void orc_xorshift_fill(int8_t *dst, size_t n_bytes, uint32_t seed);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
