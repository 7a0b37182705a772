// Shared TMA helper: cached CUtensorMap over a split-fp16 operand [rows, 2*kp] (row-major, fp16) with a
// {64 columns, box_rows} box and 128-byte swizzle (defined in gemm_tcgen05.cu).
#pragma once
#include <cuda.h>

namespace ec {
namespace tc {
int get_tensor_map(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out);
}
}  // namespace ec
