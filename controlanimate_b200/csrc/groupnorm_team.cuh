// Internal interface of the BFHWC fast-path GroupNorm kernels (groupnorm_team.cu, groupnorm_ring.cu), used by ca_groupnorm_silu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ca {

// Scratch the fast path needs for this shape (0 when the shape is outside the fast path).
size_t gn_team_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype);

// Launches the fast path when the shape fits; *handled says whether it did (otherwise the caller uses the split kernels).
int gn_team_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c, int f,
                   int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, bool* handled);

// Pipelined slice-ring kernel (groupnorm_ring.cu): same contract; tried before the team kernel.
size_t gn_ring_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype);
int gn_ring_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c, int f,
                   int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, bool* handled);

// Two streaming kernels with an L2-resident re-read (groupnorm_stream.cu): same contract; the default native-layout path.
size_t gn_stream_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype);
int gn_stream_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c,
                     int f, int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, void* workspace,
                     size_t workspace_bytes, cudaStream_t st, bool* handled);

// Small-domain kernel (groupnorm_slab.cu): a CTA owns every row of (domain, slab of whole groups); no workspace, no exchange.
int gn_slab_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c,
                   int f, int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, cudaStream_t st,
                   bool* handled);

}  // namespace ca
