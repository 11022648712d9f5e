// Internal interface of the native-layout (BFHWC) GroupNorm kernels behind ca_groupnorm_silu: the small-domain slab kernel
// (groupnorm_slab.cu) and the pipelined slice ring (groupnorm_ring.cu).  Everything they decline (fp32 storage, c % 8 != 0, the
// reference's NCFHW layout) runs on the split statistics / normalise kernels in groupnorm_silu.cu.  (Round 1's persistent
// "team" kernel and the two-launch streaming pair measured slower than these at every config-2 shape and were removed.)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ca {

// Pipelined slice-ring kernel (groupnorm_ring.cu).  gn_ring_workspace_bytes: scratch for this shape (0 when the shape is outside
// the kernel's reach); gn_ring_launch launches when the shape fits and reports through *handled whether it did.
size_t gn_ring_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype);
int gn_ring_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c, int f,
                   int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, bool* handled);

// Small-domain kernel (groupnorm_slab.cu): a CTA owns every row of (domain, slab of whole groups); no workspace, no exchange.
int gn_slab_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c,
                   int f, int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, cudaStream_t st,
                   bool* handled);

}  // namespace ca
