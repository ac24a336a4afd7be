// ssb_backward.cu -- the policy's backward kernels as their own translation unit (see ssb_backward.cuh).
#include <algorithm>

#include "ssb_backward.cuh"
#include "ssb_decima.cuh"
#include "ssb_decima_tc.cuh"

namespace ssb {
namespace bwd {
namespace {
template <int ST>
cudaError_t launch(const Params &p, int num_sms, const int32_t *list, const int32_t *offset, const int32_t *count,
                   int level, const float *g_out, float *dX, float *X_out, float *dW, Bufs b, bool many_ctas,
                   const float *x_in, cudaStream_t s)
{
    static bool prepared = false;
    if (!prepared) {
        cudaError_t e = cudaFuncSetAttribute(tc::k_mlp_backward<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)tc::BwdSmem<ST>::BYTES);
        if (e != cudaSuccess) return e;
        prepared = true;
    }
    tc::TileArgs a{list, offset, count, level};
    // as many CTAs per SM as the tile's shared memory allows (1 for the 64-wide score heads, 3 for the GNN MLPs)
    constexpr int fit = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)227 * 1024 / (tc::BwdSmem<ST>::BYTES + 1024)));
    const int per_sm = many_ctas ? fit : 1;
    const tc::BwdBufs bw{b.d_h, b.d_hdag, b.d_hglob, b.d_hinit, b.d_msg};
    tc::k_mlp_backward<ST><<<num_sms * per_sm, 128, tc::BwdSmem<ST>::BYTES, s>>>(p, a, g_out, dX, X_out, dW, bw, x_in);
    return cudaGetLastError();
}
}  // namespace

cudaError_t mlp_backward(int stage, const Params &p, int num_sms, const int32_t *list, const int32_t *offset,
                         const int32_t *count, int level, const float *g_out, float *dX, float *X_out, float *dW,
                         Bufs b, bool many_ctas, const float *x_in, cudaStream_t s)
{
    switch (stage) {
    case tc::ST_PREP: return launch<tc::ST_PREP>(p, num_sms, list, offset, count, level, g_out, dX, X_out, dW, b, many_ctas, x_in, s);
    case tc::ST_SINK: return launch<tc::ST_SINK>(p, num_sms, list, offset, count, level, g_out, dX, X_out, dW, b, many_ctas, x_in, s);
    case tc::ST_MSG: return launch<tc::ST_MSG>(p, num_sms, list, offset, count, level, g_out, dX, X_out, dW, b, many_ctas, x_in, s);
    case tc::ST_RCV: return launch<tc::ST_RCV>(p, num_sms, list, offset, count, level, g_out, dX, X_out, dW, b, many_ctas, x_in, s);
    case tc::ST_DAG: return launch<tc::ST_DAG>(p, num_sms, list, offset, count, level, g_out, dX, X_out, dW, b, many_ctas, x_in, s);
    case tc::ST_GLOB: return launch<tc::ST_GLOB>(p, num_sms, list, offset, count, level, g_out, dX, X_out, dW, b, many_ctas, x_in, s);
    case tc::ST_STAGE: return launch<tc::ST_STAGE>(p, num_sms, list, offset, count, level, g_out, dX, X_out, dW, b, many_ctas, x_in, s);
    case tc::ST_EXEC: return launch<tc::ST_EXEC>(p, num_sms, list, offset, count, level, g_out, dX, X_out, dW, b, many_ctas, x_in, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t save_rows(int stage, const Params &p, int num_sms, const int32_t *list, const int32_t *offset,
                      const int32_t *count, int level, float *x_save, cudaStream_t s)
{
    tc::TileArgs a{list, offset, count, level};
    if (stage == tc::ST_MSG) tc::k_save_rows<tc::ST_MSG><<<num_sms * 8, 128, 0, s>>>(p, a, x_save);
    else if (stage == tc::ST_RCV) tc::k_save_rows<tc::ST_RCV><<<num_sms * 8, 128, 0, s>>>(p, a, x_save);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t head_adjoint(const Params &p, const float *grad_lgprob, const float *grad_entropy, float *grad_stage,
                         float *grad_exec, cudaStream_t s)
{
    tc::k_pol_head_adjoint<<<(p.B + 3) / 4, 128, 0, s>>>(p, grad_lgprob, grad_entropy, grad_stage, grad_exec);
    return cudaGetLastError();
}

}  // namespace bwd
}  // namespace ssb
