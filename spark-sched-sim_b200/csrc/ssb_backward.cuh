// ssb_backward.cuh -- host-side launchers of the policy's backward kernels.  The kernels (ssb_decima_tc.cuh:
// k_pol_head_adjoint, k_mlp_backward<ST>) are instantiated in their own translation unit (ssb_backward.cu): the
// rollout kernels of ssb_api.cu are bound by instruction supply and their speed moves by several percent with the
// code that is laid out around them (-5.5 % with these kernels instantiated there), so they are kept out of it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ssb_types.cuh"

namespace ssb {
namespace tc {
struct BwdBufs;
}
namespace bwd {
struct Bufs { float *d_h, *d_hdag, *d_hglob, *d_hinit, *d_msg; };  // = tc::BwdBufs
// stage = tc::Stage; returns the CUDA error of the launch (cudaSuccess = ok)
cudaError_t mlp_backward(int stage, const Params &p, int num_sms, const int32_t *list, const int32_t *offset,
                         const int32_t *count, int level, const float *g_out, float *dX, float *X_out, float *dW,
                         Bufs bufs, bool many_ctas, const float *x_in, cudaStream_t s);
// the input rows of a message-passing level's list (stage = ST_MSG / ST_RCV), stored at their position in p.pl_lvl:
// what mlp_backward(x_in) reads once the embeddings they were gathered from have been overwritten
cudaError_t save_rows(int stage, const Params &p, int num_sms, const int32_t *list, const int32_t *offset,
                      const int32_t *count, int level, float *x_save, cudaStream_t s);
cudaError_t head_adjoint(const Params &p, const float *grad_lgprob, const float *grad_entropy, float *grad_stage,
                         float *grad_exec, cudaStream_t s);
}  // namespace bwd
}  // namespace ssb
