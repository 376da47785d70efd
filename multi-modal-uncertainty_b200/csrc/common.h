// Shared host-side declarations for the mmu_b200 CUDA library.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

// Error codes returned across the C ABI (0 = success).  Mirrored in include/mmu_b200.h.
#define MMU_OK 0
#define MMU_ERR_SHAPE (-1)   // unsupported / inconsistent dimensions
#define MMU_ERR_ALIGN (-2)   // pointer or leading dimension not 16-byte aligned
#define MMU_ERR_DRIVER (-3)  // CUDA driver entry point unavailable (no GPU / no libcuda)
#define MMU_ERR_TMAP (-4)    // cuTensorMapEncodeTiled rejected the descriptor
#define MMU_ERR_CUDA (-5)    // launch failed; see cudaGetLastError
#define MMU_ERR_ARG (-6)     // null pointer / bad enum
#define MMU_ERR_WORKSPACE (-7)  // workspace too small

namespace mmu {
// number of kernels this library has launched in this process (host-side counter)
void count_launch(int n = 1);
long long launch_count();
int sm_count();
// SM budget of the persistent tcgen05 GEMM grids (0 = every SM); see gemm_tcgen05.cu
void set_gemm_sm_limit(int n);
int gemm_sm_limit();
int gemm_sms();
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, long long inner, long long outer,
                      long long ld, int box_inner, int box_outer);
// 3-D (inner contiguous, mid, outer; element strides), box = box_inner x 1 x box_outer, SWIZZLE_128B
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, long long inner, long long mid,
                      long long outer, long long mid_stride, long long outer_stride, int box_inner,
                      int box_outer);
// 4-D output / epilogue-operand map (columns, rows, batch % hdiv, batch / hdiv) with a
// [32 rows x 64 B] SWIZZLE_64B box (gemm_tcgen05.cu); stores clip at the extents, loads zero-fill
int make_tmap_out_4d(CUtensorMap* out, const void* base, int is_bf16, long long cols, long long rows,
                     long long ld, long long hdiv, long long hstride, long long nmid,
                     long long mid_stride);
}  // namespace mmu
