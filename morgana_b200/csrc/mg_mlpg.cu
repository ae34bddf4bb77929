// K8 -- batched maximum-likelihood parameter generation (MLPG) on the device.
//
// Replaces viz.synthesis.MLPG (reference morgana/viz/synthesis.py:79-180), which the example models call inside predict()
// (models/RNN_SPSS.py:108-118, models/f0_test_model.py:89-90): a Python loop over batch_size x feat_dim banded solves in
// fp64 on the CPU through `bandmat`, wrapped in a device->host and a host->device copy.
//
// For one utterance and one static dimension, with L = n + 2 * padding frames (edges replicated, synthesis.py:114-120,
// 156-158) and the default windows  w0 = [1],  w1 = [-0.5, 0, 0.5],  w2 = [1, -2, 1]  (synthesis.py:122-127):
//     P = sum_k W_k^T diag(1 / var_k) W_k            (pentadiagonal, symmetric positive definite; synthesis.py:60-73)
//     b = sum_k W_k^T (mean_k / var_k)
//     solve P c = b                                   (bandmat.linalg.solveh; here an L D L^T factorisation, fp64)
// and the trajectory is c without the padding (synthesis.py:170-171).  W_k are Toeplitz band matrices truncated at the
// sequence ends (synthesis.py:8-36).
//
// One thread per (utterance, static dimension) system; neighbouring threads own neighbouring dimensions, so every load
// and store of a warp is a contiguous run.  The recurrence is sequential in time, so the work is split into a pass with no
// loop-carried dependence (build P and b, all loads in flight) and the two substitution sweeps, whose operands are
// streamed from a workspace in batches of 8 frames.  Everything is fp64, as in the reference; the output is fp32.
#include "mg_common.cuh"

namespace {

constexpr int kMlpgBatch = 8;

struct MlpgParams {
  const float* means;
  const float* variances;
  const int64_t* seq_len;
  float* out;
  double* work;        // [B][L_max][4][F]: p0 / l1, p1 / l2, p2, b / z per frame, dimension fastest
  int64_t m_sb, m_st, v_sb, v_st, o_sb, o_st, T, L_max;
  int B, F, padding;
};

__global__ void mlpg_kernel(const MlpgParams prm) {
  const int i = blockIdx.x;
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  const int F = prm.F;
  if (d >= F) return;
  int64_t n = prm.T;
  if (prm.seq_len != nullptr) {
    n = prm.seq_len[i];
    n = n < 0 ? 0 : (n > prm.T ? prm.T : n);
  }
  float* out = prm.out + i * prm.o_sb + d;
  for (int64_t t = n; t < prm.T; ++t) out[t * prm.o_st] = 0.f;     // out-of-sequence frames stay zero (synthesis.py:153)
  if (n == 0) return;
  const int pad = prm.padding;
  const int64_t L = n + 2 * pad;
  const float* mean_i = prm.means + i * prm.m_sb + d;
  const float* var_i = prm.variances + i * prm.v_sb + d;
  double* work = prm.work + static_cast<int64_t>(i) * prm.L_max * 4 * F + d;
  auto W = [&](int64_t t, int q) -> double& { return work[(t * 4 + q) * F]; };

  // precision-weighted mean and precision of window k at padded frame t (edge frames replicated)
  auto load = [&](int64_t t, int k, double& bt, double& tau) {
    int64_t tt = t - pad;
    tt = tt < 0 ? 0 : (tt > n - 1 ? n - 1 : tt);
    const double m = static_cast<double>(__ldg(mean_i + tt * prm.m_st + k * F));
    const double v = static_cast<double>(__ldg(var_i + tt * prm.v_st + k * F));
    tau = 1.0 / v;
    bt = m / v;
  };

  // ---- pass 1: P (three diagonals) and b for every frame; no loop-carried dependence.  Frames are taken four at a time:
  // the (mean, variance) pairs of frames a0-1 .. a0+4 for the three windows are all requested before any arithmetic, so the
  // loads of a batch overlap instead of paying one memory latency per frame. -----------------------------------------------
  // window coefficients at offsets (-1, 0, +1): w0 = (0, 1, 0), w1 = (-0.5, 0, 0.5), w2 = (1, -2, 1)
  constexpr int kP1 = 4;
  for (int64_t a0 = 0; a0 < L; a0 += kP1) {
    double bt[3][kP1 + 2], tau[3][kP1 + 2];     // index j <-> frame a0 - 1 + j
#pragma unroll
    for (int j = 0; j < kP1 + 2; ++j) {
      const int64_t t = a0 - 1 + j;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (t >= 0 && t < L && !(k == 0 && (j == 0 || j == kP1 + 1))) load(t, k, bt[k][j], tau[k][j]);
        else { bt[k][j] = 0.; tau[k][j] = 0.; }     // outside the sequence: the truncated window rows contribute nothing
      }
    }
#pragma unroll
    for (int u = 0; u < kP1; ++u) {
      const int64_t a = a0 + u;
      if (a >= L) break;
      const int j = u + 1;
      const bool has_next = a + 1 < L, has_next2 = a + 2 < L;
      // zeros stand in for frames outside [0, L): exactly the terms the truncated Toeplitz matrices drop
      double p0 = tau[0][j] + 4.0 * tau[2][j] + 0.25 * tau[1][j - 1] + tau[2][j - 1] + 0.25 * tau[1][j + 1] + tau[2][j + 1];
      double p1 = has_next ? -2.0 * tau[2][j] - 2.0 * tau[2][j + 1] : 0.;
      double p2 = has_next2 ? -0.25 * tau[1][j + 1] + tau[2][j + 1] : 0.;
      double bsum = bt[0][j] - 2.0 * bt[2][j] + 0.5 * bt[1][j - 1] + bt[2][j - 1] - 0.5 * bt[1][j + 1] + bt[2][j + 1];
      W(a, 0) = p0;
      W(a, 1) = p1;
      W(a, 2) = p2;
      W(a, 3) = bsum;
    }
  }

  // ---- pass 2: L D L^T factorisation + forward substitution, operands streamed in batches ------------------------------
  double d1 = 1., d2 = 1., l1_1 = 0., l2_1 = 0., l2_2 = 0., y1 = 0., y2 = 0.;   // state of frames a-1 / a-2
  for (int64_t a0 = 0; a0 < L; a0 += kMlpgBatch) {
    double p0[kMlpgBatch], p1[kMlpgBatch], p2[kMlpgBatch], bb[kMlpgBatch];
#pragma unroll
    for (int u = 0; u < kMlpgBatch; ++u) {
      const int64_t a = a0 + u < L ? a0 + u : L - 1;
      p0[u] = W(a, 0); p1[u] = W(a, 1); p2[u] = W(a, 2); bb[u] = W(a, 3);
    }
#pragma unroll
    for (int u = 0; u < kMlpgBatch; ++u) {
      const int64_t a = a0 + u;
      if (a >= L) break;
      const double da = p0[u] - l1_1 * l1_1 * d1 - l2_2 * l2_2 * d2;
      const double inv = 1.0 / da;
      const double l1 = (p1[u] - l1_1 * l2_1 * d1) * inv;
      const double l2 = p2[u] * inv;
      const double y = bb[u] - l1_1 * y1 - l2_2 * y2;
      W(a, 0) = l1;
      W(a, 1) = l2;
      W(a, 3) = y * inv;       // z = D^{-1} y
      // shift the two-frame history
      d2 = d1; d1 = da;
      l2_2 = l2_1; l2_1 = l2; l1_1 = l1;
      y2 = y1; y1 = y;
    }
  }
  // NOTE on the history variables: l1_1 = L[a][a-1] (from frame a-1), l2_2 = L[a][a-2] (from frame a-2), and the cross
  // term of l1 uses l1_1 * l2_1 with l2_1 = L[a+1][a-1] (from frame a-1).

  // ---- pass 3: back substitution c_a = z_a - l1_a c_{a+1} - l2_a c_{a+2}, again in batches -----------------------------
  double c1 = 0., c2 = 0.;
  for (int64_t a0 = L - 1; a0 >= 0; a0 -= kMlpgBatch) {
    double l1[kMlpgBatch], l2[kMlpgBatch], z[kMlpgBatch];
#pragma unroll
    for (int u = 0; u < kMlpgBatch; ++u) {
      const int64_t a = a0 - u >= 0 ? a0 - u : 0;
      l1[u] = W(a, 0); l2[u] = W(a, 1); z[u] = W(a, 3);
    }
#pragma unroll
    for (int u = 0; u < kMlpgBatch; ++u) {
      const int64_t a = a0 - u;
      if (a < 0) break;
      const double c = z[u] - l1[u] * c1 - l2[u] * c2;
      c2 = c1;
      c1 = c;
      const int64_t t = a - pad;
      if (t >= 0 && t < n) out[t * prm.o_st] = static_cast<float>(c);
    }
  }
}

}  // namespace

extern "C" int64_t mg_mlpg_workspace_bytes(int B, int64_t T, int feat_dim, int padding) {
  if (B < 0 || T < 0 || feat_dim < 0 || padding < 0) return MG_ERR_INVALID_ARG;
  return static_cast<int64_t>(B) * (T + 2 * static_cast<int64_t>(padding)) * 4 * feat_dim * static_cast<int64_t>(sizeof(double));
}

extern "C" int mg_mlpg_f32(const float* means, int64_t m_sb, int64_t m_st, const float* variances, int64_t v_sb, int64_t v_st,
                           const int64_t* seq_len, float* out, int64_t o_sb, int64_t o_st, int B, int64_t T, int feat_dim,
                           int padding, void* workspace, int64_t workspace_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && T >= 0 && feat_dim >= 0 && padding >= 0, "mg_mlpg_f32: negative shape");
  MG_REQUIRE(B <= 65535 * 32, "mg_mlpg_f32: B too large");
  if (B == 0 || T == 0 || feat_dim == 0) return MG_OK;
  MG_REQUIRE(means != nullptr && variances != nullptr && out != nullptr && workspace != nullptr, "mg_mlpg_f32: NULL buffer");
  MG_REQUIRE(workspace_bytes >= mg_mlpg_workspace_bytes(B, T, feat_dim, padding), "mg_mlpg_f32: workspace too small");
  MG_REQUIRE(mg_aligned(workspace, 8), "mg_mlpg_f32: workspace must be 8-byte aligned");
  MlpgParams prm;
  prm.means = means; prm.variances = variances; prm.seq_len = seq_len; prm.out = out;
  prm.work = static_cast<double*>(workspace);
  prm.m_sb = m_sb; prm.m_st = m_st; prm.v_sb = v_sb; prm.v_st = v_st; prm.o_sb = o_sb; prm.o_st = o_st;
  prm.T = T; prm.L_max = T + 2 * static_cast<int64_t>(padding);
  prm.B = B; prm.F = feat_dim; prm.padding = padding;
  const int threads = feat_dim >= 64 ? 64 : 32;
  dim3 grid(static_cast<unsigned>(B), static_cast<unsigned>((feat_dim + threads - 1) / threads));
  mlpg_kernel<<<grid, threads, 0, stream>>>(prm);
  MG_LAUNCH_OK();
  return MG_OK;
}
