// large_gemm.cuh -- FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) GEMM and the dense factorisation built on it.
//
// Replaces ksvd! (src/la_helper.jl:8-34, called at src/optimize.jl:291/293: "O(Nm^2)" thin SVD of Jct) by
//   G = J J'            dgemm_nt, lower tiles only (SYRK)            flops m(m+1)N, FP64 tensor pipe
//   G = L L'            blocked right-looking Cholesky (in-shared-memory diagonal blocks + DMMA panel/trailing updates)
//   L^-T (and L^-1)     blocked triangular inverse from the diagonal-block inverses, DMMA
// (the projector, the multipliers and NR's D0 = L^-1 are all expressed through L.)
//
// There is no tcgen05 FP64 MMA kind; mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4) is the FP64 tensor path on sm_100a.
// One GEMM shape serves everything:  C[M x N] (op)= A[M x K] * B[N x K]'   with A, B, C row-major ("NT").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lfpsqp {

enum GemmMode { GEMM_ASSIGN = 0, GEMM_SUB = 1, GEMM_ASSIGN_NEG = 2 };

constexpr int GM_BM = 128, GM_BK = 16, GM_LD = 20 /* BK + 4: conflict-free 64-bit fragment loads */, GM_STAGES = 3;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;  // src-size 0 => zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async16_sz(void *smem, const void *gmem, int sz) {   // sz in {0, 8, 16}: the rest is zero-filled
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Optional extras of one GEMM launch (all off by default).
//   batch > 1 : blockIdx.z selects one of `batch` independent products of identical shape (element strides batch_a/b/c);
//               used by the recursive triangular inverse, excludes split-K
//   tri bit 0 : A is upper triangular (row i is zero for k < i)  -> the K loop of a tile starts at its first row
//   tri bit 1 : B is lower triangular (row j is zero for k > j)  -> the K loop of a tile ends at its last row of B
//   nz        : zero-slab map of A and B (SKIP instantiation, SYRK of one matrix): nz[rb * nz_ld + kc] != 0 iff the 64-row block
//               rb has a non-zero in its K chunk kc (GM_BK columns); a K chunk is skipped when the A rows or the B rows of the
//               tile are all zero there (adds only +0.0 terms: the result is bit-identical unless the other operand holds Inf/NaN)
struct GemmExt {
  int64_t batch_a = 0, batch_b = 0, batch_c = 0;
  int batch = 1, tri = 0;
  const unsigned char *nz = nullptr;
  int64_t nz_ld = 0;
  int nz_rows = 0;
  // block-structure flags of the factorisation (large.cu::factorize): bf[I * bf_ld + J] != 0 iff the 64 x 64 block (I, J), I >= J,
  // of the working matrix (G, then L) may hold a non-zero.  A tile whose operands are structurally zero returns at once
  // (its contribution is exactly +0), a computed trailing tile marks its blocks as fill-in.
  //   bf_mode 1 (L21 = A21 D'):       skip iff the tile's A row blocks are zero in block column bf_col
  //   bf_mode 2 (C -= A B', A, B rows of L21): skip iff the A row blocks or the B row blocks are zero in column bf_col; else mark
  //   bf_mode 3 (recursive inverse):  skip iff the whole block range [bf_r0, bf_r1) x [bf_c0, bf_c1) (+ the batch offset) is zero
  int *bf = nullptr;
  int bf_ld = 0, bf_n = 0, bf_mode = 0, bf_a0 = 0, bf_b0 = 0, bf_col = 0, bf_r0 = 0, bf_r1 = 0, bf_c0 = 0, bf_c1 = 0, bf_batch = 0;
};

// C (op)= A * B'.  BN in {128, 64}.  256 threads = 8 warps as 2 (M) x 4 (N): warp tile 64 x (BN/4).
// lower_only: skip tiles strictly above the diagonal (SYRK / symmetric trailing update); tile (bi,bj) kept iff bi*BM+BM > bj*BN.
// ksplit > 1: split-K, slice z writes its partial into C + z*c_split_stride (ASSIGN only); the caller reduces.
// Requirements: lda, ldb even; A, B 16-byte aligned; K arbitrary (zero-filled), M, N arbitrary (predicated).
// EXT = false: the lean instantiation of the plain product (no batch / triangle bounds / block flags: X is ignored) -- the C5
// Gram runs 2 % faster without the extra prologue and its registers (166 vs 178).
template <int BN, bool SKIP = false, bool EXT = true>
__global__ void __launch_bounds__(256) dgemm_nt_kernel(int M, int N, int K, const double *__restrict__ A, int64_t lda,
                                                       const double *__restrict__ B, int64_t ldb, double *__restrict__ C,
                                                       int64_t ldc, int mode, int lower_only, int ksplit,
                                                       int64_t c_split_stride, const GemmExt X) {
  constexpr int BM = GM_BM, BK = GM_BK, LD = GM_LD, ST = GM_STAGES;
  constexpr int WN = BN / 4;       // warp tile width
  constexpr int NF = WN / 8;       // B fragments per warp per k4 step
  extern __shared__ double gsm[];
  double *As = gsm;                          // ST x BM x LD
  double *Bs = gsm + (size_t)ST * BM * LD;   // ST x BN x LD
  // tri bit 1: the K loop grows with the tile column -> walk the grid column by column from the last one, so that CTAs are issued
  // longest tile first (as the row-major order already does for tri bit 0); a mixed order left a long tile for the last wave
  int bi = blockIdx.y, bj = blockIdx.x;
  if (EXT && (X.tri & 2)) {
    const int t = blockIdx.y * gridDim.x + blockIdx.x;
    bj = (int)gridDim.x - 1 - t / (int)gridDim.y; bi = t % (int)gridDim.y;
  }
  if (lower_only && (bi * BM + BM <= bj * BN)) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (EXT && X.bf_mode) {
    if (X.bf_mode == 3) {
      const int off = (int)blockIdx.z * X.bf_batch, w = X.bf_c1 - X.bf_c0, cnt = (X.bf_r1 - X.bf_r0) * w;
      int any = 0;
      for (int e = tid; e < cnt; e += 256) any |= X.bf[(int64_t)(X.bf_r0 + off + e / w) * X.bf_ld + X.bf_c0 + off + e % w];
      if (!__syncthreads_or(any)) return;
    } else {
      const int ia = X.bf_a0 + (BM / 64) * bi, ib = X.bf_b0 + (BN / 64) * bj;
      const int fa0 = X.bf[(int64_t)ia * X.bf_ld + X.bf_col], fa1 = (ia + 1 < X.bf_n) ? X.bf[(int64_t)(ia + 1) * X.bf_ld + X.bf_col] : 0;
      if (!(fa0 | fa1)) return;
      if (X.bf_mode == 2) {
        const int fb0 = X.bf[(int64_t)ib * X.bf_ld + X.bf_col];
        const int fb1 = (BN > 64 && ib + 1 < X.bf_n) ? X.bf[(int64_t)(ib + 1) * X.bf_ld + X.bf_col] : 0;
        if (!(fb0 | fb1)) return;
        if (tid == 0) {   // fill-in (block columns > bf_col: nobody reads them during this step)
          if (fa0 & fb0) X.bf[(int64_t)ia * X.bf_ld + ib] = 1;
          if (fa1 & fb0) X.bf[(int64_t)(ia + 1) * X.bf_ld + ib] = 1;
          if (BN > 64 && (fa0 & fb1)) X.bf[(int64_t)ia * X.bf_ld + ib + 1] = 1;
          if (BN > 64 && (fa1 & fb1)) X.bf[(int64_t)(ia + 1) * X.bf_ld + ib + 1] = 1;
        }
      }
    }
  }
  const int wm = warp >> 2, wn = warp & 3;   // 2 x 4
  const int row0 = bi * BM, col0 = bj * BN;
  int zsplit = blockIdx.z;
  if (EXT && X.batch > 1) {
    A += (int64_t)blockIdx.z * X.batch_a; B += (int64_t)blockIdx.z * X.batch_b; C += (int64_t)blockIdx.z * X.batch_c;
    zsplit = 0;
  }
  // K range of this split
  int kchunks = (K + BK - 1) / BK;
  int per = (kchunks + ksplit - 1) / ksplit;
  int kc0 = zsplit * per, kc1 = min(kchunks, kc0 + per);
  if (EXT && (X.tri & 1)) kc0 = max(kc0, row0 / BK);
  if (EXT && (X.tri & 2)) kc1 = min(kc1, (col0 + BN + BK - 1) / BK);
  if (kc0 >= kc1) { kc1 = kc0; }
  // SKIP: ordered list of the K chunks of [kc0, kc1) in which both the A rows and the B rows of this tile have non-zeros
  int *klist = reinterpret_cast<int *>(gsm + (size_t)ST * (BM + BN) * LD);
  int nlist = 0;
  if (SKIP) {
    __shared__ int wcnt[8];
    const int rba = row0 / 64, rbb = col0 / 64;
    const unsigned char *ma0 = X.nz + (int64_t)rba * X.nz_ld, *ma1 = X.nz + (int64_t)min(rba + 1, X.nz_rows - 1) * X.nz_ld;
    const unsigned char *mb0 = X.nz + (int64_t)rbb * X.nz_ld, *mb1 = X.nz + (int64_t)min(rbb + (BN > 64 ? 1 : 0), X.nz_rows - 1) * X.nz_ld;
    for (int base = kc0; base < kc1; base += 256) {
      const int kc = base + (int)threadIdx.x;
      const bool f = kc < kc1 && (ma0[kc] | ma1[kc]) && (mb0[kc] | mb1[kc]);
      const unsigned bal = __ballot_sync(0xffffffffu, f);
      if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = __popc(bal);
      __syncthreads();
      int off = nlist, tot = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) { const int cw = wcnt[w]; if (w < (int)(threadIdx.x >> 5)) off += cw; tot += cw; }
      if (f) klist[off + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = kc;
      nlist += tot;
      __syncthreads();
    }
  }
  double acc[8][NF][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < NF; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  auto load_stage = [&](int stage, int kc) {
    const int k0 = kc * BK;
    // A tile: BM rows x 8 chunks of 16 B
#pragma unroll
    for (int q = 0; q < (BM * 8) / 256; q++) {
      int idx = tid + q * 256, r = idx >> 3, ch = idx & 7;
      int gr = row0 + r, gk = k0 + ch * 2;
      bool ok = (gr < M) && (gk < K);
      const double *src = ok ? (A + (int64_t)gr * lda + gk) : A;
      cp_async16_sz(As + ((size_t)stage * BM + r) * LD + ch * 2, src, ok ? (gk + 1 < K ? 16 : 8) : 0);   // odd K: never read column K
    }
#pragma unroll
    for (int q = 0; q < (BN * 8) / 256; q++) {
      int idx = tid + q * 256, r = idx >> 3, ch = idx & 7;
      int gr = col0 + r, gk = k0 + ch * 2;
      bool ok = (gr < N) && (gk < K);
      const double *src = ok ? (B + (int64_t)gr * ldb + gk) : B;
      cp_async16_sz(Bs + ((size_t)stage * BN + r) * LD + ch * 2, src, ok ? (gk + 1 < K ? 16 : 8) : 0);
    }
  };

  const int nk = SKIP ? nlist : kc1 - kc0;
  for (int s = 0; s < ST - 1; s++) {
    if (s < nk) load_stage(s, SKIP ? klist[s] : kc0 + s);
    cp_async_commit();
  }
  for (int it = 0; it < nk; it++) {
    cp_async_wait<ST - 2>();
    __syncthreads();
    {  // prefetch stage it+ST-1 (its buffer was consumed in iteration it-1)
      int nx = it + ST - 1;
      if (nx < nk) load_stage(nx % ST, SKIP ? klist[nx] : kc0 + nx);
      cp_async_commit();
    }
    const double *as = As + (size_t)(it % ST) * BM * LD + (size_t)(wm * 64) * LD;
    const double *bs = Bs + (size_t)(it % ST) * BN * LD + (size_t)(wn * WN) * LD;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      double af[8], bf[NF];
#pragma unroll
      for (int i = 0; i < 8; i++) af[i] = as[(size_t)(i * 8 + (lane >> 2)) * LD + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int j = 0; j < NF; j++) bf[j] = bs[(size_t)(j * 8 + (lane >> 2)) * LD + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < NF; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  // epilogue: thread holds C[row = lane/4][col = 2*(lane%4) + {0,1}] of each 8x8 atom
  double *Cz = C + (int64_t)zsplit * c_split_stride;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int r = row0 + wm * 64 + i * 8 + (lane >> 2);
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < NF; j++) {
      int c = col0 + wn * WN + j * 8 + 2 * (lane & 3);
      double *p = Cz + (int64_t)r * ldc + c;
      if (c + 1 < N) {
        double2 v;
        if (mode == GEMM_ASSIGN) v = make_double2(acc[i][j][0], acc[i][j][1]);
        else if (mode == GEMM_ASSIGN_NEG) v = make_double2(-acc[i][j][0], -acc[i][j][1]);
        else { double2 o = *reinterpret_cast<double2 *>(p); v = make_double2(o.x - acc[i][j][0], o.y - acc[i][j][1]); }
        *reinterpret_cast<double2 *>(p) = v;
      } else if (c < N) {
        if (mode == GEMM_ASSIGN) p[0] = acc[i][j][0];
        else if (mode == GEMM_ASSIGN_NEG) p[0] = -acc[i][j][0];
        else p[0] -= acc[i][j][0];
      }
    }
  }
}

template <int BN> constexpr size_t dgemm_smem_bytes() { return (size_t)GM_STAGES * (GM_BM + BN) * GM_LD * sizeof(double); }

// Zero-slab map of a row-major matrix (rows x cols, ld): nz[rb * nz_ld + kc] = 1 iff rows [64 rb, 64 rb + 64) hold a non-zero
// (or NaN) in columns [GM_BK kc, GM_BK kc + GM_BK).  One pass over the matrix at HBM speed; grid (ceil(cols / 1024), ceil(rows / 64)).
// count[0] += number of non-zero map entries (the host reads the density with the control block).
__global__ void __launch_bounds__(256) zero_slab_map_kernel(const double *__restrict__ A, int64_t ld, int rows, int64_t cols,
                                                            unsigned char *__restrict__ nz, int64_t nz_ld, unsigned long long *count) {
  __shared__ int flag[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 64) flag[tid] = 0;
  __syncthreads();
  const int r0 = blockIdx.y * 64;
  const int64_t c0 = (int64_t)blockIdx.x * 1024;
  unsigned mine = 0;   // bit s: chunk 4 s + lane / 8 of this CTA's 64 chunks
  for (int rr = warp; rr < 64; rr += 8) {
    const int r = r0 + rr;
    if (r >= rows) break;
    const double *row = A + (int64_t)r * ld + c0;
#pragma unroll 4
    for (int s = 0; s < 16; s++) {
      const int64_t c = (int64_t)s * 64 + 2 * lane;
      if (c0 + c + 1 < cols) {
        const double2 v = *reinterpret_cast<const double2 *>(row + c);
        if (v.x != 0.0 || v.y != 0.0) mine |= 1u << s;
      } else if (c0 + c < cols) {
        if (row[c] != 0.0) mine |= 1u << s;
      }
    }
  }
  for (int s = 0; s < 16; s++) if (mine >> s & 1u) flag[4 * s + (lane >> 3)] = 1;
  __syncthreads();
  if (tid < 64) {
    const int64_t kc = (int64_t)blockIdx.x * 64 + tid;
    const int f = flag[tid];
    if (kc < nz_ld) nz[(int64_t)blockIdx.y * nz_ld + kc] = (unsigned char)f;
    const unsigned bal = __ballot_sync(0xffffffffu, f != 0);
    if (lane == 0 && bal) atomicAdd(count, (unsigned long long)__popc(bal));
  }
}

// ------------------------------------------------------------------ diagonal block: in-shared-memory Cholesky + inverse
// One CTA (256 threads as a 16 x 16 grid, each owning a 4 x 4 register tile) factors the 64 x 64 diagonal block of A
// (row-major, lda): right-looking Cholesky with the current column broadcast through shared memory (2 barriers per
// column), then D = L_jj^-1 by a blocked triangular inversion (16 x 16 diagonal blocks by forward substitution,
// off-diagonal blocks by two small products per block).  Writes L_jj back (zeros above the diagonal) and D
// (row-major NB x NB, zeros above the diagonal and outside nb_act) to Dout.
// flag[0] is set to 1 when a pivot is not > thresh (rank deficiency, cf. optimize.jl:297-302).
template <int NB>
__device__ __forceinline__ void potf2_inv_body(double *A, int64_t lda, int nb_act, double *Dout, const double *thresh_p, int *flag) {
  static_assert(NB == 64, "register tiling assumes a 64 x 64 block");
  extern __shared__ double psm[];
  double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(psm);
  double (*Ds)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(psm + NB * (NB + 1));
  __shared__ double colbuf[NB];
  __shared__ double Ts[16][17];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const double thresh = thresh_p[0];
  double a[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      int gr = 4 * ty + r, gc = 4 * tx + c;
      a[r][c] = (gr < nb_act && gc < nb_act) ? A[(int64_t)gr * lda + gc] : (gr == gc ? 1.0 : 0.0);
    }
  for (int k = 0; k < NB; k++) {
    const int kb = k >> 2, kk = k & 3;
    if (tx == kb) {   // owners of column k publish it (unscaled)
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) if (c == kk) colbuf[4 * ty + r] = a[r][c];
    }
    __syncthreads();
    double piv = colbuf[k];
    if (!(piv > thresh)) { if (tid == 0 && k < nb_act) flag[0] = 1; piv = (piv > 0.0) ? piv : 1.0; }
    const double d = sqrt(piv), dinv = 1.0 / d;   // correctly rounded on purpose: rsqrt-based pivots (1-2 ulp) moved bound-active solves by 1e-7
    double lr[4], lc[4];
#pragma unroll
    for (int r = 0; r < 4; r++) { int gr = 4 * ty + r; lr[r] = (gr > k) ? colbuf[gr] * dinv : 0.0; }
#pragma unroll
    for (int c = 0; c < 4; c++) { int gc = 4 * tx + c; lc[c] = (gc > k) ? colbuf[gc] * dinv : 0.0; }
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) a[r][c] -= lr[r] * lc[c];
    if (tx == kb) {   // store the scaled column into the owners' registers
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) if (c == kk) { int gr = 4 * ty + r; a[r][c] = (gr > k) ? lr[r] : (gr == k ? d : a[r][c]); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      int gr = 4 * ty + r, gc = 4 * tx + c;
      Ls[gr][gc] = (gc <= gr) ? a[r][c] : 0.0;
      Ds[gr][gc] = 0.0;
    }
  __syncthreads();
  // phase 1: the four 16 x 16 diagonal blocks, one column per thread (64 threads)
  if (tid < 64) {
    const int b = tid >> 4, c = tid & 15, o = 16 * b;
    for (int i = c; i < 16; i++) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int t = c; t < i; t++) s -= Ls[o + i][o + t] * Ds[o + t][o + c];
      Ds[o + i][o + c] = s / Ls[o + i][o + i];
    }
  }
  __syncthreads();
  // phase 2: off-diagonal blocks by increasing block distance: D_ij = -D_ii * (sum_{k=j}^{i-1} L_ik D_kj)
  for (int dist = 1; dist < 4; dist++) {
    for (int j = 0; j + dist < 4; j++) {
      const int i = j + dist;
      double t = 0.0;
      for (int kb2 = j; kb2 < i; kb2++)
#pragma unroll
        for (int q = 0; q < 16; q++) t += Ls[16 * i + ty][16 * kb2 + q] * Ds[16 * kb2 + q][16 * j + tx];
      Ts[ty][tx] = t;
      __syncthreads();
      double u = 0.0;
#pragma unroll
      for (int q = 0; q < 16; q++) u += Ds[16 * i + ty][16 * i + q] * Ts[q][tx];
      __syncthreads();
      Ds[16 * i + ty][16 * j + tx] = -u;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) {
    int r = e / NB, c = e % NB;
    if (r < nb_act && c < nb_act) A[(int64_t)r * lda + c] = Ls[r][c];
    Dout[e] = (r < nb_act && c < nb_act) ? Ds[r][c] : 0.0;
  }
}
// chain call of the blocked Cholesky; done[b] != 0: the block was already factorised by potf2_inv_indep_kernel
template <int NB>
__global__ void __launch_bounds__(256) potf2_inv_kernel(double *A, int64_t lda, int nb_act, double *Dout,
                                                        const double *thresh_p, int *flag, const int *done = nullptr, int b = 0) {
  if (done && done[b]) return;
  potf2_inv_body<NB>(A, lda, nb_act, Dout, thresh_p, flag);
}
// Diagonal blocks with no structural non-zero to their left (bf row b, columns < b) never receive an update: they are
// factorised here, all at once (one CTA per block), before the chain starts.  Thomson: G is diagonal -> all 64 blocks.
template <int NB>
__global__ void __launch_bounds__(256) potf2_inv_indep_kernel(double *G, int64_t ldg, int m, double *Dblk, const double *thresh_p,
                                                              int *flag, const int *bf, int nblk, int *done, int *dependent) {
  const int b = blockIdx.x;
  int any = 0;
  for (int J = threadIdx.x; J < b; J += 256) any |= bf[(int64_t)b * nblk + J];
  any = __syncthreads_or(any);
  if (any) { if (threadIdx.x == 0) { done[b] = 0; atomicAdd(dependent, 1); } return; }
  potf2_inv_body<NB>(G + (int64_t)b * NB * ldg + (int64_t)b * NB, ldg, min(NB, m - b * NB), Dblk + (size_t)b * NB * NB, thresh_p, flag);
  if (threadIdx.x == 0) done[b] = 1;
}

// bf[I * ld + J] = 1 iff the 64 x 64 block (I, J), I >= J, of the lower triangle of G holds a non-zero (or NaN); grid (nblk, nblk)
__global__ void __launch_bounds__(256) block_nz_kernel(const double *__restrict__ G, int64_t ldg, int m, int *bf, int ld) {
  const int I = blockIdx.y, J = blockIdx.x;
  if (J > I) { if (threadIdx.x == 0) bf[(int64_t)I * ld + J] = 0; return; }
  int any = 0;
  for (int e = threadIdx.x; e < 64 * 32; e += 256) {
    const int r = 64 * I + e / 32, c = 64 * J + 2 * (e % 32);
    if (r < m && c < m) {
      const double2 v = *reinterpret_cast<const double2 *>(G + (int64_t)r * ldg + c);   // the pad column (odd m) is never non-zero
      any |= (v.x != 0.0) | ((c + 1 < m) & (v.y != 0.0));
    }
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) bf[(int64_t)I * ld + J] = any ? 1 : 0;
}

// thresh[0] = max(eps_rank^2, 1e-14 * max_i G[i][i])  (the Gram-form stand-in for "sigma_j < eps_rank", optimize.jl:297-302)
__global__ void __launch_bounds__(256) diag_thresh_kernel(const double *G, int64_t ldg, int m, double eps_rank, double *thresh) {
  __shared__ double red[256];
  double mx = 0.0;
  for (int i = threadIdx.x; i < m; i += 256) mx = fmax(mx, G[(int64_t)i * ldg + i]);
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
  if (threadIdx.x == 0) thresh[0] = fmax(eps_rank * eps_rank, 1e-14 * red[0]);
}

// dst (cols x rows, ldd) = src (rows x cols, lds)'  -- 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_kernel(const double *__restrict__ src, int64_t lds, double *__restrict__ dst,
                                                        int64_t ldd, int rows, int cols, int64_t batch_src = 0, int64_t batch_dst = 0) {
  __shared__ double t[32][33];
  src += (int64_t)blockIdx.z * batch_src; dst += (int64_t)blockIdx.z * batch_dst;
  int bx = blockIdx.x * 32, by = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    int gr = by + r, gc = bx + tx;
    t[r][tx] = (gr < rows && gc < cols) ? src[(int64_t)gr * lds + gc] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int gr = bx + r, gc = by + tx;  // dst row = src col
    if (gr < cols && gc < rows) dst[(int64_t)gr * ldd + gc] = t[tx][r];
  }
}

// zero the strict upper triangle of a row-major m x m matrix / copy a small block / set to zero
__global__ void zero_upper_kernel(double *A, int64_t lda, int m) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)m * m) return;
  int r = (int)(e / m), c = (int)(e % m);
  if (c > r) A[(int64_t)r * lda + c] = 0.0;
}
// dst[c][r] = src[r][c] for r, c < nb_act ; src is an NB x NB row-major block
__global__ void copy_block_T_kernel(const double *src, int NB, int nb_act, double *dst, int64_t ldd) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NB * NB) return;
  int r = e / NB, c = e % NB;
  if (r < nb_act && c < nb_act) dst[(int64_t)c * ldd + r] = src[e];
}
// level 0 of the recursive triangular inverse: diagonal block b of XT = D_b', of Linv = D_b (D_b = L_bb^-1, NB x NB, zero-padded)
__global__ void __launch_bounds__(256) diag_blocks_kernel(const double *Dblk, int NB, int m, double *XT, double *Linv, int64_t ld) {
  const int b = blockIdx.x, i0 = b * NB, nb = min(NB, m - i0);
  const double *D = Dblk + (size_t)b * NB * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    if (r < nb && c < nb) {
      const double v = D[e];
      Linv[(int64_t)(i0 + r) * ld + i0 + c] = v;
      XT[(int64_t)(i0 + c) * ld + i0 + r] = v;
    }
  }
}

}  // namespace lfpsqp
