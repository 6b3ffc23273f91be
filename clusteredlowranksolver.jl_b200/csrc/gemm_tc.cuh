// gemm_tc.cuh — tcgen05 int8 slice-pair GEMM (filled in below the CUDA-core path).
#pragma once
