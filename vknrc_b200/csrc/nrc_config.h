// nrc_config.h -- network shape, hyper-parameters and buffer layouts of the NRC hot path (host + device).
// Values and layouts are the reference's (paths relative to the VkNRC tree):
//   shape        src/VkNRCState.hpp:22-25        batches      shader/src/Constant.glsl:6-7
//   Adam / EMA   shader/src/Constant.glsl:10-13, shader/src/nrc_optimize.comp:10-11
//   records      shader/src/NRCRecord.glsl:6-38, src/VkNRCState.cpp:10-32
#pragma once
#include <stdint.h>

#define NRC_WIDTH 64
#define NRC_OUT_WIDTH 3
#define NRC_HIDDEN_LAYERS 5
#define NRC_LAYERS 6
#define NRC_WEIGHT_COUNT (NRC_WIDTH * NRC_WIDTH * NRC_HIDDEN_LAYERS + NRC_WIDTH * NRC_OUT_WIDTH) /* 20672 */
#define NRC_WEIGHT_ROWS (NRC_WEIGHT_COUNT / NRC_WIDTH)                                          /* 323 */
#define NRC_TRAIN_BATCH_COUNT 4
#define NRC_TRAIN_BATCH_SIZE 16384
#define NRC_TILE 128 /* samples per MMA tile == the reference's workgroup (NN_nv.glsl:12-14) */

#define NRC_LOSS_SCALE 1.0f
#define NRC_ADAM_BETA1 0.9f
#define NRC_ADAM_BETA2 0.999f
#define NRC_EMA_ALPHA 0.99f
#define NRC_LEARNING_RATE 0.002f
#define NRC_ADAM_EPSILON 1e-8f
#define NRC_EVAL_INVALID_DST 0xFFFFFFFFu

/* `gradients` buffer: [0, 20672) dW in the weight layout (fp32), then two bookkeeping slots so that ONE
 * all-reduce carries everything a replicated optimizer step needs. Padded to a multiple of 64 floats. */
#define NRC_GRAD_LOSS_SLOT NRC_WEIGHT_COUNT       /* sum over records of the per-record loss */
#define NRC_GRAD_COUNT_SLOT (NRC_WEIGHT_COUNT + 1) /* number of records that contributed (as float, exact < 2^24) */
#define NRC_GRAD_STRIDE (NRC_WEIGHT_COUNT + 64)    /* 20736 */

#include "../../include/nrc_b200_types.h" /* record / optimizer / scene structs shared with the public C ABI */
