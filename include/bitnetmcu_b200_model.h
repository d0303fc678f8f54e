/*
 * bitnetmcu_b200_model.h -- plain-C glue: build the runtime layer table of include/bitnetmcu_b200.h from the
 * compile-time macros and arrays of a generated BitNetMCU_model.h.
 *
 * The reference selects its model by #include "BitNetMCU_model.h" and hard-codes the layer names in
 * BitMnistInference (/root/reference/BitNetMCU_MNIST_dll.c:48-121: L1..L4 for MODEL_FCMNIST with the optional
 * L4 behind #ifdef L4_active, L2/L4/L6/L7/L9/L11/L13/L15 for MODEL_CNNMNIST).  This header uses exactly the same
 * names, so any header the reference's own dll.c compiles against also works here.  Include it AFTER
 * BitNetMCU_model.h and bitnetmcu_b200.h:
 *
 *     static const bnm_layer layers[] = { BNM_MODEL_LAYERS };
 *     bnm_model_create(BNM_MODEL_CLASS, layers, sizeof layers / sizeof layers[0], 256, device, &model);
 */
#ifndef BITNETMCU_B200_MODEL_H
#define BITNETMCU_B200_MODEL_H

#define BNM_FC_LAYER(k) \
    { BNM_LAYER_FC, L##k##_bitperweight, L##k##_incoming_weights, L##k##_outgoing_weights, 0, 0, L##k##_weights, sizeof(L##k##_weights) }
#define BNM_CONV_LAYER(k) \
    { BNM_LAYER_CONV33, L##k##_bitperweight, L##k##_incoming_x, L##k##_out_channels, L##k##_in_channels, L##k##_groups, L##k##_weights, sizeof(L##k##_weights) }
#define BNM_POOL_LAYER(k) \
    { BNM_LAYER_MAXPOOL22, 0, L##k##_incoming_x, L##k##_outgoing_x, 0, 0, 0, 0 }

#if defined(MODEL_CNNMNIST)
#define BNM_MODEL_CLASS BNM_MODEL_CNNMNIST
/* dll.c:64-90 */
#define BNM_MODEL_LAYERS \
    BNM_CONV_LAYER(2), BNM_CONV_LAYER(4), BNM_POOL_LAYER(6), BNM_CONV_LAYER(7), BNM_POOL_LAYER(9), \
    BNM_FC_LAYER(11), BNM_FC_LAYER(13), BNM_FC_LAYER(15)
#define BNM_MODEL_N_CLASSES L15_outgoing_weights
#elif defined(MODEL_FCMNIST)
#define BNM_MODEL_CLASS BNM_MODEL_FCMNIST
/* dll.c:95-121 */
#ifdef L4_active
#define BNM_MODEL_LAYERS BNM_FC_LAYER(1), BNM_FC_LAYER(2), BNM_FC_LAYER(3), BNM_FC_LAYER(4)
#define BNM_MODEL_N_CLASSES L4_outgoing_weights
#else
#define BNM_MODEL_LAYERS BNM_FC_LAYER(1), BNM_FC_LAYER(2), BNM_FC_LAYER(3)
#define BNM_MODEL_N_CLASSES L3_outgoing_weights
#endif
#else
#error "No model defined"   /* same diagnostic as BitNetMCU_MNIST_dll.c:122-124 */
#endif

#endif
