// Internal declarations shared by the .cu translation units of libviditq_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/viditq_b200.h"

namespace vq {
int num_sms();
int make_u8_kmajor_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch,
                        uint32_t box_rows);
}  // namespace vq
