"""vknrc_b200 -- B200-native (sm_100a) implementation of VkNRC's one data-parallel hot path: the Neural Radiance
Cache's fully-fused 64-wide MLP (query inference + online training). The product is ``libnrc_b200.so`` (CUDA kernels
+ the C ABI of ``include/nrc_b200.h``); this package only binds it for tests, bench and examples."""
from .api import (EVAL_RECORD_DTYPE, GRAD_COUNT_SLOT, GRAD_LOSS_SLOT, GRADIENT_FLOATS, MATERIAL_DTYPE, TRAIN_BATCH_COUNT,
                  TRAIN_BATCH_SIZE, TRAIN_RECORD_DTYPE, WEIGHT_COUNT, DeviceScene, NrcError, NrcState, lib, mlp_evaluate_encoded,
                  mlp_gradient_encoded, unpack_inputs, encode_inputs, encode_packed_inputs)

__all__ = ["NrcState", "NrcError", "lib", "mlp_evaluate_encoded", "mlp_gradient_encoded", "WEIGHT_COUNT", "GRADIENT_FLOATS",
           "GRAD_LOSS_SLOT", "GRAD_COUNT_SLOT", "TRAIN_BATCH_SIZE", "TRAIN_BATCH_COUNT", "DeviceScene", "unpack_inputs", "encode_inputs", "encode_packed_inputs", "MATERIAL_DTYPE",
           "EVAL_RECORD_DTYPE", "TRAIN_RECORD_DTYPE"]
