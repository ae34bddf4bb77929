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
// The recurrences of a banded solve are sequential in time, and a batch holds only batch_size x feat_dim systems (32 for a
// 1-dimensional log-F0 stream), so one thread per system leaves the GPU idle and pays one memory + arithmetic latency per
// frame (2.6 ms for 1 600 frames).  Here time is parallel too:
//
//   build  one thread per (utterance, frame, dimension): the three diagonals of P and b, fully parallel, coalesced reads.
//   solve  one WARP per system.  The L frames are cut into S <= 32 chunks separated by 2-frame separators (bandwidth 2, so
//          chunk interiors only couple to their adjacent separators).  Lane j factorises its interior block A_j (L D L^T) and
//          solves A_j [g | U_L | U_R] = [b_j | B_L | B_R] -- its right-hand side and the four coupling columns -- in one
//          forward and one backward sweep of ~L/32 steps.  The separators then satisfy a block-tridiagonal SPD system with
//          2 x 2 blocks (the Schur complement), assembled from the first / last two rows of g, U_L, U_R by neighbour
//          shuffles and solved by block elimination along the warp (S - 1 short steps).  Finally every lane forms
//          c_j = g_j - U_L x_{j-1} - U_R x_j for its frames, with no dependence between frames.
//          This is exact block elimination in a nested-dissection order (no truncation), fp64 throughout as in the
//          reference; the output is fp32.  ~4 L / 32 + 2 S dependent steps instead of 2 L.
#include <string.h>

#include "mg_common.cuh"

namespace {

constexpr int kMaxWindows = 4;
constexpr int kWork = 8;            // doubles per (system, frame) in the workspace: one 64-byte line
constexpr int kSolveWarps = 4;      // systems per CTA of the solve kernel
constexpr int kAhead = 4;           // frames whose operands are requested before the dependent arithmetic of a batch
constexpr int kMinChunk = 12;       // frames per chunk (interior + separator) below which a sequence is not split further

struct MlpgParams {
  const float* means;
  const float* variances;
  const int64_t* seq_len;
  float* out;
  double* work;        // [B * F systems][L_max frames][kWork]
  int64_t m_sb, m_st, v_sb, v_st, o_sb, o_st, T, L_max;
  int B, F, padding;
  int n_windows;            // 1 .. kMaxWindows windows of extent <= 1 frame on either side
  double coef[4][3];        // window k at offsets (-1, 0, +1)
};

__device__ __forceinline__ int64_t mlpg_valid(const MlpgParams& prm, int i) {
  int64_t n = prm.T;
  if (prm.seq_len != nullptr) {
    n = prm.seq_len[i];
    n = n < 0 ? 0 : (n > prm.T ? prm.T : n);
  }
  return n;
}

// ---- build: P (three diagonals, upper storage: p0 = P[a][a], p1 = P[a][a+1], p2 = P[a][a+2]) and b per padded frame ----
__global__ void __launch_bounds__(256) mlpg_build_kernel(const MlpgParams prm) {
  const int F = prm.F;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(prm.B) * prm.L_max * F;
  if (idx >= total) return;
  const int d = static_cast<int>(idx % F);
  const int64_t a = (idx / F) % prm.L_max;
  const int i = static_cast<int>(idx / (static_cast<int64_t>(F) * prm.L_max));
  const int64_t n = mlpg_valid(prm, i);
  if (a < prm.T && a >= n) prm.out[i * prm.o_sb + a * prm.o_st + d] = 0.f;   // out-of-sequence frames stay zero (synthesis.py:153)
  const int pad = prm.padding;
  const int64_t L = n + 2 * pad;
  if (n == 0 || a >= L) return;
  const float* mean_i = prm.means + i * prm.m_sb + d;
  const float* var_i = prm.variances + i * prm.v_sb + d;
  // precision-weighted mean and precision of window k at padded frame t (edge frames replicated); zero outside [0, L):
  // exactly the terms the truncated Toeplitz window matrices drop
  auto load = [&](int64_t t, int k, double& bt, double& tau) {
    if (t < 0 || t >= L) { bt = 0.; tau = 0.; return; }
    int64_t tt = t - pad;
    tt = tt < 0 ? 0 : (tt > n - 1 ? n - 1 : tt);
    const double m = static_cast<double>(__ldg(mean_i + tt * prm.m_st + k * F));
    const double v = static_cast<double>(__ldg(var_i + tt * prm.v_st + k * F));
    tau = 1.0 / v;
    bt = m * tau;     // one fp64 division per operand pair (within 1 ulp of m / v)
  };
  // Window k has coefficients c_k at offsets (-1, 0, +1) (defaults: (0, 1, 0), (-0.5, 0, 0.5), (1, -2, 1)); with W_k[t][t + j] =
  // c_k[j] truncated at the ends:
  //   P[a][a]   = sum_k sum_{t = a-1..a+1} c_k[a - t]^2 tau_k[t]
  //   P[a][a+1] = sum_k c_k[0] c_k[+1] tau_k[a] + c_k[-1] c_k[0] tau_k[a+1]
  //   P[a][a+2] = sum_k c_k[-1] c_k[+1] tau_k[a+1]
  //   b[a]      = sum_k sum_{t = a-1..a+1} c_k[a - t] (mean_k / var_k)[t]
  // Zero coefficients are skipped (warp-uniform), so the default windows cost what they did before.
  double p0 = 0., p1 = 0., p2 = 0., bsum = 0.;
  for (int k = 0; k < prm.n_windows; ++k) {
    const double cm = prm.coef[k][0], c0 = prm.coef[k][1], cp = prm.coef[k][2];
    double bt, tau;
    if (c0 != 0.) {
      load(a, k, bt, tau);
      p0 += c0 * c0 * tau;
      p1 += c0 * cp * tau;
      bsum += c0 * bt;
    }
    if (cp != 0.) {            // frame a - 1 reaches a through its +1 coefficient
      load(a - 1, k, bt, tau);
      p0 += cp * cp * tau;
      bsum += cp * bt;
    }
    if (cm != 0.) {            // frame a + 1 reaches a through its -1 coefficient
      load(a + 1, k, bt, tau);
      p0 += cm * cm * tau;
      p1 += cm * c0 * tau;
      p2 += cm * cp * tau;
      bsum += cm * bt;
    }
  }
  if (a + 1 >= L) p1 = 0.;
  if (a + 2 >= L) p2 = 0.;
  double2* w = reinterpret_cast<double2*>(prm.work + ((static_cast<int64_t>(i) * F + d) * prm.L_max + a) * kWork);
  w[0] = make_double2(p0, p1);
  w[1] = make_double2(p2, bsum);
}

// The five solution vectors a lane carries for one row of its chunk: A^-1 b and the four coupling columns.
struct Row5 {
  double g, a, b, r0, r1;   // g | U_L (separator frames s-2, s-1) | U_R (separator frames e, e+1)
};

__global__ void __launch_bounds__(kSolveWarps * 32) mlpg_solve_kernel(const MlpgParams prm) {
  const int lane = threadIdx.x & 31;
  const int64_t sys = static_cast<int64_t>(blockIdx.x) * kSolveWarps + (threadIdx.x >> 5);
  const int F = prm.F;
  if (sys >= static_cast<int64_t>(prm.B) * F) return;          // warp-uniform
  const int i = static_cast<int>(sys / F), d = static_cast<int>(sys % F);
  const int64_t n = mlpg_valid(prm, i);
  if (n == 0) return;
  const int pad = prm.padding;
  const int64_t L = n + 2 * pad;
  double* work = prm.work + sys * prm.L_max * kWork;
  float* out = prm.out + i * prm.o_sb + d;

  // ---- chunk geometry (warp-uniform S; lane j < S owns chunk j = interior frames [start, start + m), followed by the
  // separator frames start + m and start + m + 1 unless it is the last chunk) ------------------------------------------
  const int S = L >= 2 * kMinChunk ? static_cast<int>(min(static_cast<int64_t>(32), L / kMinChunk)) : 1;
  const int64_t base = L - 2 * (S - 1);
  const int64_t m_lo = base / S, rem = base % S;
  const bool active = lane < S;
  const int64_t m = active ? m_lo + (lane < rem ? 1 : 0) : 0;
  const int64_t start = lane * m_lo + min(static_cast<int64_t>(lane), rem) + 2 * lane;
  const bool has_left = active && lane > 0, has_right = active && lane < S - 1;
  auto line = [&](int64_t frame) { return reinterpret_cast<double2*>(work + frame * kWork); };

  // coupling coefficients: B_L rows 0, 1 against separator frames (s-2, s-1); B_R rows m-2, m-1 against (e, e+1)
  double cL00 = 0., cL10 = 0., cL11 = 0., cR00 = 0., cR01 = 0., cR11 = 0.;
  if (has_left) {
    const double2 q2 = line(start - 2)[1], q1a = line(start - 1)[0], q1b = line(start - 1)[1];
    cL00 = q2.x;    // P[s-2][s]   -> row 0, separator frame s-2
    cL10 = q1a.y;   // P[s-1][s]   -> row 0, separator frame s-1
    cL11 = q1b.x;   // P[s-1][s+1] -> row 1, separator frame s-1
  }

  // ---- forward sweep: L D L^T of the interior block and forward substitution of g, U_L (U_R's columns are zero above the
  // last two rows, so their forward substitution is three values kept in registers) -------------------------------------
  double zR0a = 0., zR0b = 0., zR1b = 0.;
  {
    double d1 = 1., d2 = 1., l1_1 = 0., l2_1 = 0., l2_2 = 0.;       // factor history of rows r-1 / r-2
    double g1 = 0., g2 = 0., a1 = 0., a2 = 0., b1 = 0., b2 = 0., yR0a = 0.;
    // Software pipeline: the operands of batch k + 1 are requested BEFORE batch k is computed and stored (the stores go
    // to the same array, so the compiler may not move later loads above them), which hides the memory latency of a batch
    // under the dependent arithmetic of the previous one.
    double2 pa[kAhead], pb[kAhead], na[kAhead], nb[kAhead];
    auto fetch = [&](int64_t q0, double2* fa, double2* fb) {
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {
        const int64_t r = q0 + u < m ? q0 + u : m - 1;
        fa[u] = line(start + r)[0];
        fb[u] = line(start + r)[1];
      }
    };
    if (m > 0) fetch(0, pa, pb);
    for (int64_t q0 = 0; q0 < m; q0 += kAhead) {
      if (q0 + kAhead < m) fetch(q0 + kAhead, na, nb);
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {
        const int64_t r = q0 + u;
        if (r >= m) break;
        const double p0 = pa[u].x, bb = pb[u].y;
        double p1 = pa[u].y, p2 = pb[u].x;
        if (has_right) {     // entries that reach into the right separator belong to B_R, not to the interior block
          if (r == m - 1) { cR01 = p1; cR11 = p2; p1 = 0.; p2 = 0.; }
          else if (r == m - 2) { cR00 = p2; p2 = 0.; }
        }
        const double da = p0 - l1_1 * l1_1 * d1 - l2_2 * l2_2 * d2;
        const double inv = 1.0 / da;
        const double l1 = (p1 - l1_1 * l2_1 * d1) * inv;   // L[r+1][r]
        const double l2 = p2 * inv;                        // L[r+2][r]
        const double rhs_a = r == 0 ? cL00 : 0.;
        const double rhs_b = r == 0 ? cL10 : (r == 1 ? cL11 : 0.);
        const double yg = bb - l1_1 * g1 - l2_2 * g2;
        const double ya = rhs_a - l1_1 * a1 - l2_2 * a2;
        const double yb = rhs_b - l1_1 * b1 - l2_2 * b2;
        if (has_right) {
          if (r == m - 2) { yR0a = cR00; zR0a = yR0a * inv; }
          else if (r == m - 1) { zR0b = (cR01 - l1_1 * yR0a) * inv; zR1b = cR11 * inv; }
        }
        double2* w = line(start + r);
        w[0] = make_double2(l1, l2);
        w[1] = make_double2(yg * inv, ya * inv);            // z = D^-1 y
        w[2] = make_double2(yb * inv, 0.);
        d2 = d1; d1 = da;
        l2_2 = l2_1; l2_1 = l2; l1_1 = l1;
        g2 = g1; g1 = yg; a2 = a1; a1 = ya; b2 = b1; b1 = yb;
      }
#pragma unroll
      for (int u = 0; u < kAhead; ++u) { pa[u] = na[u]; pb[u] = nb[u]; }
    }
  }

  // ---- backward sweep: c_r = z_r - l1_r c_{r+1} - l2_r c_{r+2} for the five vectors; the solutions replace the factors in
  // the workspace, and the first / last two rows stay in registers for the separator system -----------------------------
  Row5 F0 = {0., 0., 0., 0., 0.}, F1 = F0, E0 = F0, E1 = F0;
  {
    Row5 c1 = {0., 0., 0., 0., 0.}, c2 = c1;
    double2 wa[kAhead], wb[kAhead], wc[kAhead], xa[kAhead], xb[kAhead], xc[kAhead];
    auto fetch = [&](int64_t q0, double2* fa, double2* fb, double2* fc) {
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {
        const int64_t r = q0 - u >= 0 ? q0 - u : 0;
        fa[u] = line(start + r)[0];
        fb[u] = line(start + r)[1];
        fc[u] = line(start + r)[2];
      }
    };
    if (m > 0) fetch(m - 1, wa, wb, wc);
    for (int64_t q0 = m - 1; q0 >= 0; q0 -= kAhead) {
      if (q0 - kAhead >= 0) fetch(q0 - kAhead, xa, xb, xc);
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {
        const int64_t r = q0 - u;
        if (r < 0) break;
        const double l1 = wa[u].x, l2 = wa[u].y;
        const double zr0 = has_right ? (r == m - 1 ? zR0b : (r == m - 2 ? zR0a : 0.)) : 0.;
        const double zr1 = has_right && r == m - 1 ? zR1b : 0.;
        Row5 c;
        c.g = wb[u].x - l1 * c1.g - l2 * c2.g;
        c.a = wb[u].y - l1 * c1.a - l2 * c2.a;
        c.b = wc[u].x - l1 * c1.b - l2 * c2.b;
        c.r0 = zr0 - l1 * c1.r0 - l2 * c2.r0;
        c.r1 = zr1 - l1 * c1.r1 - l2 * c2.r1;
        double2* w = line(start + r);
        w[0] = make_double2(c.g, c.a);
        w[1] = make_double2(c.b, c.r0);
        w[2] = make_double2(c.r1, 0.);
        if (r == m - 1) E1 = c;
        if (r == m - 2) E0 = c;
        if (r == 1) F1 = c;
        if (r == 0) F0 = c;
        c2 = c1;
        c1 = c;
      }
#pragma unroll
      for (int u = 0; u < kAhead; ++u) { wa[u] = xa[u]; wb[u] = xb[u]; wc[u] = xc[u]; }
    }
  }

  // ---- the separators' system  M_s x_s + Lo_s x_{s-1} + Up_s x_{s+1} = r_s,  s = 0 .. S-2 (lane s owns separator s,
  // the two frames after its chunk) ----------------------------------------------------------------------------------------
  double x0 = 0., x1 = 0.;          // solution of this lane's separator
  if (S > 1) {                      // warp-uniform
    // contributions of this lane's chunk: (B^T v) for v in {g, U_L columns, U_R columns}
    auto BR0 = [&](double v_m2, double v_m1) { return cR00 * v_m2 + cR01 * v_m1; };   // separator frame e
    auto BR1 = [&](double v_m1) { return cR11 * v_m1; };                              // separator frame e + 1
    auto BL0 = [&](double v_0) { return cL00 * v_0; };                                // separator frame s - 2
    auto BL1 = [&](double v_0, double v_1) { return cL10 * v_0 + cL11 * v_1; };       // separator frame s - 1
    // left side (goes to the previous lane's separator)
    double elg0 = BL0(F0.g), elg1 = BL1(F0.g, F1.g);
    double ell00 = BL0(F0.a), ell10 = BL1(F0.a, F1.a), ell01 = BL0(F0.b), ell11 = BL1(F0.b, F1.b);      // (B_L^T U_L)
    double elr00 = BL0(F0.r0), elr10 = BL1(F0.r0, F1.r0), elr01 = BL0(F0.r1), elr11 = BL1(F0.r1, F1.r1);  // (B_L^T U_R)
    // fetch the next chunk's left-side terms
    elg0 = __shfl_down_sync(MG_FULL_MASK, elg0, 1); elg1 = __shfl_down_sync(MG_FULL_MASK, elg1, 1);
    ell00 = __shfl_down_sync(MG_FULL_MASK, ell00, 1); ell10 = __shfl_down_sync(MG_FULL_MASK, ell10, 1);
    ell01 = __shfl_down_sync(MG_FULL_MASK, ell01, 1); ell11 = __shfl_down_sync(MG_FULL_MASK, ell11, 1);
    elr00 = __shfl_down_sync(MG_FULL_MASK, elr00, 1); elr10 = __shfl_down_sync(MG_FULL_MASK, elr10, 1);
    elr01 = __shfl_down_sync(MG_FULL_MASK, elr01, 1); elr11 = __shfl_down_sync(MG_FULL_MASK, elr11, 1);
    double m00 = 1., m01 = 0., m10 = 0., m11 = 1., lo00 = 0., lo01 = 0., lo10 = 0., lo11 = 0.;
    double up00 = 0., up01 = 0., up10 = 0., up11 = 0., r0 = 0., r1 = 0.;
    if (has_right) {
      const int64_t e = start + m;
      const double2 de0 = line(e)[0], de1 = line(e)[1], df0 = line(e + 1)[0], df1 = line(e + 1)[1];
      // D = [[P[e][e], P[e][e+1]], [P[e][e+1], P[e+1][e+1]]],  f = (b[e], b[e+1])
      m00 = de0.x - BR0(E0.r0, E1.r0) - ell00;
      m01 = de0.y - BR0(E0.r1, E1.r1) - ell01;
      m10 = de0.y - BR1(E1.r0) - ell10;
      m11 = df0.x - BR1(E1.r1) - ell11;
      lo00 = -BR0(E0.a, E1.a); lo01 = -BR0(E0.b, E1.b);
      lo10 = -BR1(E1.a);       lo11 = -BR1(E1.b);
      up00 = -elr00; up01 = -elr01; up10 = -elr10; up11 = -elr11;
      r0 = de1.y - BR0(E0.g, E1.g) - elg0;
      r1 = df1.y - BR1(E1.g) - elg1;
    }
    // block elimination along the warp: G_s = M'_s^-1 Up_s, h_s = M'_s^-1 r'_s
    double G00 = 0., G01 = 0., G10 = 0., G11 = 0., h0 = 0., h1 = 0.;
    for (int s = 0; s < S - 1; ++s) {
      const int src = s > 0 ? s - 1 : 0;
      const double pG00 = __shfl_sync(MG_FULL_MASK, G00, src), pG01 = __shfl_sync(MG_FULL_MASK, G01, src);
      const double pG10 = __shfl_sync(MG_FULL_MASK, G10, src), pG11 = __shfl_sync(MG_FULL_MASK, G11, src);
      const double ph0 = __shfl_sync(MG_FULL_MASK, h0, src), ph1 = __shfl_sync(MG_FULL_MASK, h1, src);
      if (lane == s) {
        double a00 = m00, a01 = m01, a10 = m10, a11 = m11, q0 = r0, q1 = r1;
        if (s > 0) {
          a00 -= lo00 * pG00 + lo01 * pG10; a01 -= lo00 * pG01 + lo01 * pG11;
          a10 -= lo10 * pG00 + lo11 * pG10; a11 -= lo10 * pG01 + lo11 * pG11;
          q0 -= lo00 * ph0 + lo01 * ph1;    q1 -= lo10 * ph0 + lo11 * ph1;
        }
        const double idet = 1.0 / (a00 * a11 - a01 * a10);
        const double i00 = a11 * idet, i01 = -a01 * idet, i10 = -a10 * idet, i11 = a00 * idet;
        G00 = i00 * up00 + i01 * up10; G01 = i00 * up01 + i01 * up11;
        G10 = i10 * up00 + i11 * up10; G11 = i10 * up01 + i11 * up11;
        h0 = i00 * q0 + i01 * q1;      h1 = i10 * q0 + i11 * q1;
      }
    }
    for (int s = S - 2; s >= 0; --s) {
      const int src = s + 1 < 32 ? s + 1 : 31;
      const double n0 = __shfl_sync(MG_FULL_MASK, x0, src), n1 = __shfl_sync(MG_FULL_MASK, x1, src);
      if (lane == s) {
        x0 = h0; x1 = h1;
        if (s < S - 2) { x0 -= G00 * n0 + G01 * n1; x1 -= G10 * n0 + G11 * n1; }
      }
    }
  }

  // ---- every frame of the chunk: c = g - U_L x_left - U_R x_right (no dependence between frames) -----------------------
  const double xl0 = __shfl_up_sync(MG_FULL_MASK, x0, 1), xl1 = __shfl_up_sync(MG_FULL_MASK, x1, 1);
  const double kl0 = has_left ? xl0 : 0., kl1 = has_left ? xl1 : 0., kr0 = has_right ? x0 : 0., kr1 = has_right ? x1 : 0.;
  const int64_t r_lo = max(static_cast<int64_t>(0), pad - start), r_hi = min(m, n + pad - start);   // rows inside [pad, pad + n)
  for (int64_t q0 = r_lo; q0 < r_hi; q0 += kAhead) {
    double2 wa[kAhead], wb[kAhead], wc[kAhead];
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      const int64_t r = q0 + u < r_hi ? q0 + u : r_hi - 1;
      wa[u] = line(start + r)[0];
      wb[u] = line(start + r)[1];
      wc[u] = line(start + r)[2];
    }
#pragma unroll
    for (int u = 0; u < kAhead; ++u) {
      const int64_t r = q0 + u;
      if (r >= r_hi) break;
      const double c = wa[u].x - wa[u].y * kl0 - wb[u].x * kl1 - wb[u].y * kr0 - wc[u].x * kr1;
      out[(start + r - pad) * prm.o_st] = static_cast<float>(c);
    }
  }
  if (has_right) {
    const int64_t t0 = start + m - pad;
    if (t0 >= 0 && t0 < n) out[t0 * prm.o_st] = static_cast<float>(x0);
    if (t0 + 1 >= 0 && t0 + 1 < n) out[(t0 + 1) * prm.o_st] = static_cast<float>(x1);
  }
}

}  // namespace

extern "C" int64_t mg_mlpg_workspace_bytes(int B, int64_t T, int feat_dim, int padding) {
  if (B < 0 || T < 0 || feat_dim < 0 || padding < 0) return MG_ERR_INVALID_ARG;
  return static_cast<int64_t>(B) * (T + 2 * static_cast<int64_t>(padding)) * kWork * feat_dim * static_cast<int64_t>(sizeof(double));
}

extern "C" int mg_mlpg_f32(const float* means, int64_t m_sb, int64_t m_st, const float* variances, int64_t v_sb, int64_t v_st,
                           const int64_t* seq_len, float* out, int64_t o_sb, int64_t o_st, int B, int64_t T, int feat_dim,
                           int padding, const double* windows, int n_windows, void* workspace, int64_t workspace_bytes,
                           mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && T >= 0 && feat_dim >= 0 && padding >= 0, "mg_mlpg_f32: negative shape");
  MG_REQUIRE(B <= 65535 * 32, "mg_mlpg_f32: B too large");
  if (B == 0 || T == 0 || feat_dim == 0) return MG_OK;
  MG_REQUIRE(means != nullptr && variances != nullptr && out != nullptr && workspace != nullptr, "mg_mlpg_f32: NULL buffer");
  MG_REQUIRE(workspace_bytes >= mg_mlpg_workspace_bytes(B, T, feat_dim, padding), "mg_mlpg_f32: workspace too small");
  MG_REQUIRE(mg_aligned(workspace, 16), "mg_mlpg_f32: workspace must be 16-byte aligned");
  MlpgParams prm;
  prm.means = means; prm.variances = variances; prm.seq_len = seq_len; prm.out = out;
  prm.work = static_cast<double*>(workspace);
  prm.m_sb = m_sb; prm.m_st = m_st; prm.v_sb = v_sb; prm.v_st = v_st; prm.o_sb = o_sb; prm.o_st = o_st;
  prm.T = T; prm.L_max = T + 2 * static_cast<int64_t>(padding);
  prm.B = B; prm.F = feat_dim; prm.padding = padding;
  if (windows == nullptr) {      // the reference's defaults (synthesis.py:122-127)
    static const double defaults[3][3] = {{0., 1., 0.}, {-0.5, 0., 0.5}, {1., -2., 1.}};
    prm.n_windows = 3;
    memcpy(prm.coef, defaults, sizeof(defaults));
  } else {
    MG_REQUIRE(n_windows >= 1 && n_windows <= kMaxWindows, "mg_mlpg_f32: %d windows (1 .. %d are provided)", n_windows, kMaxWindows);
    prm.n_windows = n_windows;
    memcpy(prm.coef, windows, sizeof(double) * 3 * n_windows);
  }
  const int64_t build_threads = static_cast<int64_t>(B) * prm.L_max * feat_dim;
  const int64_t build_ctas = (build_threads + 255) / 256;
  const int64_t systems = static_cast<int64_t>(B) * feat_dim;
  MG_REQUIRE(build_ctas < (int64_t(1) << 31) && systems < (int64_t(1) << 31), "mg_mlpg_f32: problem too large");
  mlpg_build_kernel<<<static_cast<unsigned>(build_ctas), 256, 0, stream>>>(prm);
  MG_LAUNCH_OK();
  mlpg_solve_kernel<<<static_cast<unsigned>((systems + kSolveWarps - 1) / kSolveWarps), kSolveWarps * 32, 0, stream>>>(prm);
  MG_LAUNCH_OK();
  return MG_OK;
}
