// losses.KLD_standard_normal (reference morgana/losses.py:64-67):
//   kld  = -0.5 * sum_{last dim} (1 + log_variance - mean ** 2 - exp(log_variance))      one value per row
//   loss = mean over rows
// Forward: every CTA owns a fixed, contiguous chunk of the flat tensors, adds its terms in a fixed order (fp64 per thread,
// fixed shuffle tree, warps in index order) and leaves one partial; a one-CTA finisher adds the partials in index order and
// rounds once.  Backward (the same pass, when asked for): d loss / d mean = mean / rows, d loss / d log_variance =
// 0.5 * (exp(log_variance) - 1) / rows, both scaled by the upstream gradient read from device memory (no host sync).
#include "mg_common.cuh"

namespace {

constexpr int kKldThreads = 256;
constexpr int64_t kKldChunk = 16384;   // elements per CTA

template <bool GRAD>
__global__ void __launch_bounds__(kKldThreads)
kld_kernel(const float* __restrict__ mean, const float* __restrict__ log_var, int64_t n, double inv_rows,
           const float* __restrict__ grad_scale_dev, float* __restrict__ grad_mean, float* __restrict__ grad_log_var,
           double* __restrict__ partial) {
  __shared__ double s_warp[kKldThreads / 32];
  const int64_t begin = static_cast<int64_t>(blockIdx.x) * kKldChunk;
  const int64_t end = min(n, begin + kKldChunk);
  float w = 0.f;
  if (GRAD) w = static_cast<float>(static_cast<double>(grad_scale_dev != nullptr ? __ldg(grad_scale_dev) : 1.f) * inv_rows);
  double acc = 0.;
  for (int64_t i = begin + threadIdx.x; i < end; i += kKldThreads) {
    const float m = __ldcs(mean + i), lv = __ldcs(log_var + i);
    const float e = expf(lv);                                  // full-precision expf, as torch.exp
    // ((1 + lv) - m * m) - exp(lv): the reference's left-to-right fp32 evaluation
    acc += static_cast<double>(__fsub_rn(__fsub_rn(__fadd_rn(1.f, lv), __fmul_rn(m, m)), e));
    if (GRAD) {
      grad_mean[i] = __fmul_rn(m, w);
      grad_log_var[i] = __fmul_rn(__fmul_rn(0.5f, __fsub_rn(e, 1.f)), w);
    }
  }
  acc = mg_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.;
#pragma unroll
    for (int i = 0; i < kKldThreads / 32; ++i) s += s_warp[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256)
kld_finish_kernel(const double* __restrict__ partial, int n_partials, double inv_rows, float* __restrict__ loss) {
  __shared__ double s_warp[8];
  double acc = 0.;
  for (int i = threadIdx.x; i < n_partials; i += 256) acc += partial[i];
  acc = mg_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += s_warp[i];
    *loss = static_cast<float>(-0.5 * s * inv_rows);
  }
}

int kld_ctas(int64_t n) { return static_cast<int>((n + kKldChunk - 1) / kKldChunk); }

}  // namespace

extern "C" int64_t mg_kld_workspace_bytes(int64_t n) { return n > 0 ? static_cast<int64_t>(kld_ctas(n)) * 8 : 8; }

extern "C" int mg_kld_standard_normal_f32(const float* mean, const float* log_variance, int64_t rows, int latent_dim, float* loss,
                                          float* grad_mean, float* grad_log_variance, const float* grad_scale_dev,
                                          void* workspace, int64_t workspace_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(rows >= 1 && latent_dim >= 1, "mg_kld_standard_normal_f32: bad shape (rows=%lld, latent_dim=%d)", static_cast<long long>(rows), latent_dim);
  const int64_t n = rows * latent_dim;
  MG_REQUIRE(n / latent_dim == rows && kld_ctas(n) < (1 << 30), "mg_kld_standard_normal_f32: too many elements");
  MG_REQUIRE(mean != nullptr && log_variance != nullptr, "mg_kld_standard_normal_f32: NULL operand");
  MG_REQUIRE((grad_mean == nullptr) == (grad_log_variance == nullptr), "mg_kld_standard_normal_f32: both gradients or none");
  MG_REQUIRE(loss != nullptr || grad_mean != nullptr, "mg_kld_standard_normal_f32: nothing to compute");
  MG_REQUIRE(workspace != nullptr && workspace_bytes >= mg_kld_workspace_bytes(n) && mg_aligned(workspace, 8),
             "mg_kld_standard_normal_f32: workspace of %lld bytes needed", static_cast<long long>(mg_kld_workspace_bytes(n)));
  const int n_ctas = kld_ctas(n);
  const double inv_rows = 1. / static_cast<double>(rows);
  double* partial = static_cast<double*>(workspace);
  if (grad_mean != nullptr)
    kld_kernel<true><<<n_ctas, kKldThreads, 0, stream>>>(mean, log_variance, n, inv_rows, grad_scale_dev, grad_mean, grad_log_variance, partial);
  else
    kld_kernel<false><<<n_ctas, kKldThreads, 0, stream>>>(mean, log_variance, n, inv_rows, nullptr, nullptr, nullptr, partial);
  MG_LAUNCH_OK();
  if (loss != nullptr) {
    kld_finish_kernel<<<1, 256, 0, stream>>>(partial, n_ctas, inv_rows, loss);
    MG_LAUNCH_OK();
  }
  return MG_OK;
}
