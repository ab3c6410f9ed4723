// large_eig.cu -- rank-deficient Jacobians in large-n mode: eigen-decomposition of the replicated m x m Gram and its truncated
// pseudo-inverse (the Gram-form equivalent of the reference's rank scan on the singular values, src/optimize.jl:297-302, of
// kgemv! on the leading `rank` columns, src/la_helper.jl:36-44, and of the zeroed multipliers, optimize.jl:335-340).
//
//   G = J W J' = V diag(lambda) V'            one-sided (Hestenes) Jacobi, cyclic round-robin ordering, on the device
//   rank = #{lambda_k >= max(eps_rank^2, 1e-13 lambda_max)}        (sigma_k = sqrt(lambda_k); same rule as batched mode)
//   G^+  = V_r diag(1 / lambda_r) V_r'        replaces (L L')^-1 in every solve:  U_r U_r' = J' G^+ J,  lambda = G^+ J (-g)
//
// One-sided Jacobi works on the ROWS of Bt = (G V)' and Vt = V' (row-major, contiguous): a rotation of the column pair (p, q)
// of G V and V is a rotation of the rows p, q of Bt and Vt.  Rotations are chosen to make the rows of Bt orthogonal; at
// convergence |Bt_k| = lambda_k and Vt_k is the eigenvector.  The m/2 disjoint pairs of one round-robin step are independent
// (one CTA per pair, pairs distributed over the grid), m-1 steps make a sweep, a cooperative grid barrier separates steps.
// It runs only when the Cholesky pivot test of the factorisation fails (or after it failed once for this problem): a
// fallback path, 8-12 sweeps.  Every rank of a column-sharded solve holds the same all-reduced Gram and runs it redundantly.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <algorithm>

#include "large_device.cuh"
#include "large_state.h"

using namespace lfpsqp;

namespace {
constexpr int ET = 256;

struct EigBar {   // grid barrier of the cooperative launch (as large_fused.cu)
  unsigned *ctr; unsigned gen, nblk;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      gen++;
      const unsigned target = gen * nblk;
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
    }
    __syncthreads();
  }
};

// Bt (in: the symmetric Gram, full; out: rows = lambda_k v_k), Vt (out: rows = eigenvectors), conv[sweep] = max |cos angle|
__global__ void __launch_bounds__(ET) jacobi_eig_kernel(double *Bt, double *Vt, int64_t ld, int m, int max_sweeps, double tol,
                                                        unsigned long long *conv, unsigned *bar, int *sweeps_done) {
  __shared__ double sh[3 * 33];
  EigBar grid{bar, 0u, gridDim.x};
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = (m + 1) & ~1;                 // players of the round-robin tournament (one dummy when m is odd)
  const int half = M / 2;
  for (int i = blockIdx.x; i < m; i += gridDim.x)      // Vt = I
    for (int k = tid; k < m; k += ET) Vt[(int64_t)i * ld + k] = (i == k) ? 1.0 : 0.0;
  grid.sync();
  int sweep = 0;
  for (; sweep < max_sweeps; sweep++) {
    for (int step = 0; step < M - 1; step++) {
      for (int k = blockIdx.x; k < half; k += gridDim.x) {
        int p, q;
        if (k == 0) { p = M - 1; q = step; }
        else { p = (step + k) % (M - 1); q = (step - k + (M - 1)) % (M - 1); }
        if (p >= m || q >= m) continue;       // the dummy player
        if (p > q) { const int t = p; p = q; q = t; }
        double *bp = Bt + (int64_t)p * ld, *bq = Bt + (int64_t)q * ld;
        double a = 0.0, b = 0.0, g = 0.0;
        for (int i = tid; i < m; i += ET) { const double x = bp[i], y = bq[i]; a += x * x; b += y * y; g += x * y; }
        a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
        __syncthreads();
        if (lane == 0) { sh[warp] = a; sh[33 + warp] = b; sh[66 + warp] = g; }
        __syncthreads();
        a = 0.0; b = 0.0; g = 0.0;
        for (int w = 0; w < ET / 32; w++) { a += sh[w]; b += sh[33 + w]; g += sh[66 + w]; }   // same order in every thread
        const double den = sqrt(a * b);
        if (!(den > 0.0)) continue;           // a zero row: an exact null direction already
        const double cosang = fabs(g) / den;
        if (tid == 0) atomicMax(conv + sweep, (unsigned long long)__double_as_longlong(cosang));
        if (cosang <= 1e-17) continue;
        const double zeta = (b - a) / (2.0 * g);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        double *vp = Vt + (int64_t)p * ld, *vq = Vt + (int64_t)q * ld;
        for (int i = tid; i < m; i += ET) {
          const double x = bp[i], y = bq[i];
          bp[i] = c * x - s * y; bq[i] = s * x + c * y;
          const double u = vp[i], v = vq[i];
          vp[i] = c * u - s * v; vq[i] = s * u + c * v;
        }
      }
      grid.sync();
    }
    // all CTAs read the same value behind the barrier
    const double worst = __longlong_as_double((long long)*reinterpret_cast<volatile unsigned long long *>(conv + sweep));
    if (worst < tol) { sweep++; break; }
  }
  if (blockIdx.x == 0 && tid == 0) *sweeps_done = sweep;
}

// lambda_k = |Bt_k| ; threshold ; Vs_k = Vt_k / sqrt(lambda_k) (kept) or 0 (dropped) ; rank and lambda_max to ctrl
__global__ void __launch_bounds__(ET) eig_scale_kernel(const double *Bt, double *Vt, int64_t ld, int m, double eps_rank, double *lam, LargeCtrl *ctrl) {
  __shared__ double sh[33];
  __shared__ double s_thr;
  __shared__ int s_rank;
  // pass 1: eigenvalues (one row per warp round) and their maximum
  double lmax = 0.0;
  for (int k = threadIdx.x >> 5; k < m; k += ET / 32) {
    double s = 0.0;
    for (int i = threadIdx.x & 31; i < m; i += 32) { const double x = Bt[(int64_t)k * ld + i]; s += x * x; }
    s = sqrt(warp_sum(s));
    if ((threadIdx.x & 31) == 0) lam[k] = s;
    lmax = fmax(lmax, s);
  }
  lmax = block_max(lmax, sh);
  if (threadIdx.x == 0) { s_thr = fmax(eps_rank * eps_rank, 1e-13 * lmax); s_rank = 0; }
  __syncthreads();
  int cnt = 0;
  for (int k = threadIdx.x >> 5; k < m; k += ET / 32) {
    const double l = lam[k];
    const bool keep = l >= s_thr;
    const double sc = keep ? 1.0 / sqrt(l) : 0.0;
    for (int i = threadIdx.x & 31; i < m; i += 32) Vt[(int64_t)k * ld + i] *= sc;
    if ((threadIdx.x & 31) == 0 && keep) cnt++;
  }
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_rank, cnt);
  __syncthreads();
  if (threadIdx.x == 0) { ctrl->s[14] = (double)s_rank; ctrl->s[15] = lmax; }
}

__global__ void symmetrize_kernel(double *G, int64_t ld, int m) {   // upper triangle <- lower triangle
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)m * m) return;
  const int r = (int)(e / m), c = (int)(e % m);
  if (c > r) G[(int64_t)r * ld + c] = G[(int64_t)c * ld + r];
}
}  // namespace

// G (full symmetric, m x m, leading dimension ld; destroyed) -> Vs (rows = eigenvectors scaled by 1/sqrt(lambda), zero rows for the
// dropped ones) in `Vt`; ctrl->s[14] = rank, ctrl->s[15] = lambda_max.  `lower_only`: G holds only its lower triangle.
// Returns 0, or 1 when the cooperative launch is not possible.
int large_eig_pinv_factors(LargeState &S, double *G, double *Vt, double *lam, int lower_only, int *sweeps_dev, unsigned long long *conv_dev,
                           unsigned *bar_dev) {
  const int m = S.m; const int64_t ld = S.ldm;
  if (lower_only) symmetrize_kernel<<<(unsigned)(((int64_t)m * m + 255) / 256), 256, 0, S.stream>>>(G, ld, m);
  cudaMemsetAsync(conv_dev, 0, 32 * sizeof(unsigned long long), S.stream);
  cudaMemsetAsync(bar_dev, 0, sizeof(unsigned), S.stream);
  int grid = std::min(S.sm_count * 2, std::max(1, (m + 1) / 2));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_eig_kernel, ET, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return 1; }
  grid = std::min(grid, per_sm * S.sm_count);
  int max_sweeps = 30; double tol = 1e-15;
  int mm = m;
  void *args[] = {&G, &Vt, (void *)&ld, &mm, &max_sweeps, &tol, &conv_dev, &bar_dev, &sweeps_dev};
  if (cudaLaunchCooperativeKernel((void *)jacobi_eig_kernel, dim3(grid), dim3(ET), args, 0, S.stream) != cudaSuccess) { cudaGetLastError(); return 1; }
  eig_scale_kernel<<<1, ET, 0, S.stream>>>(G, Vt, ld, m, S.prm.eps_rank, lam, S.ctrl);
  S.launches += 3;
  return 0;
}
