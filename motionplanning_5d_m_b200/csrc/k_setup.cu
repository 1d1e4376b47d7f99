// k_setup.cu -- one-off device set-up per (H, weights): the shared Gram operator of the QP.
//
// Every inequality row CFS_FANUC.get_con (Lib/CFS_FANUC.m:119-129) and the quadprog bounds (:85) produce is a
// combination of at most nj rows of P = [B_theta; B_omega; I] (3n x n), where B_theta/B_omega are the position /
// velocity rows of sys_info.Baug (main_FANUC.m:84-86; closed form (0.5+(i-j))dt^2 and dt for j<=i).
// With QQ = L L' shared by the whole batch, G = P QQ^{-1} P' = Y'Y, Y = L^{-1} P', is computed ONCE here; the
// per-problem dual active-set solver (k_qp.cu) then never touches an n-vector: every inner product
// c_a QQ^{-1} c_b' it needs is a <=nj x nj bilinear form over entries of G (4.5 MB at n=250: L2 resident).
#include "cfs_kernels.cuh"

namespace cfs {

// ---- right-looking Cholesky of an n x n matrix held in global memory (column-major, lower), one CTA ----------
__global__ void __launch_bounds__(1024) k_chol(int n, double *A, int *info) {
  __shared__ double piv;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      const double d = A[j + (size_t)n * j];
      if (!(d > 0.0)) bad = j + 1;
      piv = sqrt(d);
    }
    __syncthreads();
    if (bad) break;
    const double d = piv;
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) A[i + (size_t)n * j] = (i == j) ? d : A[i + (size_t)n * j] / d;
    __syncthreads();
    // trailing update: A[i][c] -= L[i][j]*L[c][j] for j < c <= i
    const int m = n - j - 1;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int c = j + 1 + e / m, i = j + 1 + e % m;
      if (i >= c) A[i + (size_t)n * c] -= A[i + (size_t)n * j] * A[c + (size_t)n * j];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *info = bad;
}

// b_pi = P'(:,pi): column pi of P' = row pi of P, as a function of the control index (j,k)
__device__ __forceinline__ double p_entry(int pi, int col, int n, int nj, double dt) {
  const int j = col / nj, k = col % nj;  // control u_{j,k}, j = 0..H-1
  if (pi < n) {                          // theta primitive (i,kk): (0.5 + (i-j)) dt^2 for j <= i
    const int i = pi / nj, kk = pi % nj;
    return (kk == k && j <= i) ? (0.5 * dt * dt + ((i - j) * dt) * dt) : 0.0;
  } else if (pi < 2 * n) {               // omega primitive: dt for j <= i
    const int q = pi - n, i = q / nj, kk = q % nj;
    return (kk == k && j <= i) ? dt : 0.0;
  }
  return (pi - 2 * n == col) ? 1.0 : 0.0;  // the control itself (bounds)
}

// Yt[pi + 3n*i] = (L^{-1} P')(i, pi): one thread per primitive pi, forward substitution, pi is the fast index so
// every load of Yt is coalesced and every load of L is a warp-uniform broadcast.
__global__ void k_trsm_primitives(int n, int nj, double dt, const double *L /*or nullptr = identity*/, double *Yt) {
  const int np = 3 * n;
  const int pi = blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= np) return;
  for (int i = 0; i < n; ++i) {
    double s = p_entry(pi, i, n, nj, dt);
    if (L) {
      const double *Li = L + i;  // L[i + n*k]
#pragma unroll 4
      for (int k = 0; k < i; ++k) s -= Li[(size_t)n * k] * Yt[pi + (size_t)np * k];
      s /= Li[(size_t)n * i];
    }
    Yt[pi + (size_t)np * i] = s;
  }
}

// ---- small tiled FP64 GEMM with arbitrary strides: C[m + ldc*nn] = alpha * sum_k A(m,k) B(k,nn) ----------------
#define GT 64
#define GK 16
__global__ void __launch_bounds__(256) k_dgemm(int M, int N, int K, double alpha, const double *A, long long sAm,
                                               long long sAk, const double *B, long long sBk, long long sBn, double *C,
                                               int ldc) {
  __shared__ double As[GK][GT + 1];
  __shared__ double Bs[GK][GT + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      const int mm = e % GT, kk = e / GT;
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[gm * sAm + gk * sAk] : 0.0;
      const int gn = n0 + mm;
      Bs[kk][mm] = (gn < N && gk < K) ? B[gk * sBk + gn * sBn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[kk][tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][ty + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gm = m0 + tx + 16 * a, gn = n0 + ty + 16 * b;
      if (gm < M && gn < N) C[gm + (size_t)ldc * gn] = alpha * acc[a][b];
    }
}

cudaError_t launch_dgemm(int M, int N, int K, double alpha, const double *A, int lda, bool transA, const double *B,
                         int ldb, double *C, int ldc, cudaStream_t s) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((M + GT - 1) / GT, (N + GT - 1) / GT);
  const long long sAm = transA ? lda : 1, sAk = transA ? 1 : lda;
  k_dgemm<<<grid, 256, 0, s>>>(M, N, K, alpha, A, sAm, sAk, B, 1, ldb, C, ldc);
  return cudaGetLastError();
}

__global__ void k_diag(int np, const double *G, double *gdiag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) gdiag[i] = sqrt(G[i + (size_t)np * i]);  // QQ^-1 norm of each primitive row
}

__global__ void k_copy(size_t cnt, const double *src, double *dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// ---- N2: cost matrix and per-problem linear terms on the device ----------------------------------------------------------
// One thread per entry of QQ.  Column/row index (j,k): control of joint k at step j.  With b(i,j) = (0.5+(i-j))dt^2:
//   QQ[(j1,k1),(j2,k2)] = sum_{i >= max(j1,j2)} w_i * ( b1 b2 Qtt + b1 dt Qtw + dt b2 Qwt + dt^2 Qww )[k1,k2]  + r_scale (Rblk+Rblk')[k1,k2] [j1==j2]
__global__ void k_build_qq(int H, int nj, double dt, const double *Q, const double *Rblk, double r_scale, double stage_w,
                           double term_w, double *QQ) {
  const int n = H * nj, ns = 2 * nj;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * n) return;
  const int r = e % n, c = e / n;
  const int j1 = r / nj, k1 = r % nj, j2 = c / nj, k2 = c % nj;
  const double qtt = Q[k1 + ns * k2], qtw = Q[k1 + ns * (nj + k2)], qwt = Q[(nj + k1) + ns * k2], qww = Q[(nj + k1) + ns * (nj + k2)];
  double acc = 0.0;
  for (int i = (j1 > j2 ? j1 : j2); i < H; ++i) {
    const double w = (i == H - 1) ? term_w : stage_w;
    const double b1 = 0.5 * dt * dt + ((i - j1) * dt) * dt, b2 = 0.5 * dt * dt + ((i - j2) * dt) * dt;
    acc += w * (((b1 * b2) * qtt + (b1 * dt) * qtw) + ((dt * b2) * qwt + (dt * dt) * qww));
  }
  if (j1 == j2) acc += r_scale * (Rblk[k1 + nj * k2] + Rblk[k2 + nj * k1]);
  QQ[r + (size_t)n * c] = acc;
}

cudaError_t launch_build_qq(int H, int nj, double dt, const double *Q, const double *Rblk, double r_scale, double stage_w,
                            double term_w, double *QQ, cudaStream_t s) {
  const int n = H * nj;
  k_build_qq<<<(n * n + 255) / 256, 256, 0, s>>>(H, nj, dt, Q, Rblk, r_scale, stage_w, term_w, QQ);
  return cudaGetLastError();
}

// one CTA per problem; q_i = w_i Q e_i with e_i = [theta0 - thetag ; 0] (x0 has zero velocity, so Aaug x0 = x0 at every step)
__global__ void __launch_bounds__(128) k_build_problems(int B, int H, int nj, double dt, const double *Q, double stage_w,
                                                        double term_w, const double *theta0, const double *thetag, double *x0,
                                                        double *xref /*nullptr: keep the caller's reference*/, double *ff,
                                                        double *caug) {
  __shared__ double e[CFS_MAXL], qe_t[CFS_MAXL], qe_w[CFS_MAXL];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = H * nj, ns = 2 * nj;
  if (tid < nj) e[tid] = theta0[(size_t)b * nj + tid] - thetag[(size_t)b * nj + tid];
  __syncthreads();
  if (tid < ns) {  // (Q e)(tid) with e = [e_theta; 0]
    double acc = 0.0;
    for (int k = 0; k < nj; ++k) acc += Q[tid + ns * k] * e[k];
    if (tid < nj) qe_t[tid] = acc; else qe_w[tid - nj] = acc;
  }
  __syncthreads();
  if (tid < ns) x0[(size_t)b * ns + tid] = tid < nj ? theta0[(size_t)b * nj + tid] : 0.0;
  // x_ : MATLAB linspace(theta0, thetag, H+1) without its first column, zero velocity rows (main_FANUC.m:38-49)
  for (int idx = tid; xref && idx < H * ns; idx += blockDim.x) {
    const int i = idx / ns, k = idx % ns;
    double v = 0.0;
    if (k < nj) {
      const double t0 = theta0[(size_t)b * nj + k], tg = thetag[(size_t)b * nj + k];
      v = (i == H - 1) ? tg : t0 + ((i + 1) * (tg - t0)) / H;
    }
    xref[(size_t)b * H * ns + idx] = v;
  }
  // ff_(j,k) = sum_{i>=j} w_i ( b(i,j) (Qe)_theta(k) + dt (Qe)_omega(k) )
  for (int idx = tid; idx < n; idx += blockDim.x) {
    const int j = idx / nj, k = idx % nj;
    double acc = 0.0;
    for (int i = j; i < H; ++i) {
      const double w = (i == H - 1) ? term_w : stage_w;
      acc += w * ((0.5 * dt * dt + ((i - j) * dt) * dt) * qe_t[k] + dt * qe_w[k]);
    }
    ff[(size_t)b * n + idx] = acc;
  }
  if (tid == 0) {  // caug = sum_i w_i e'Qe
    double eq = 0.0;
    for (int k = 0; k < nj; ++k) eq += e[k] * qe_t[k];
    caug[b] = ((H - 1) * stage_w + term_w) * eq;
  }
}

cudaError_t launch_build_problems(int B, int H, int nj, double dt, const double *Q, double stage_w, double term_w,
                                  const double *theta0, const double *thetag, double *x0, double *xref, double *ff,
                                  double *caug, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  k_build_problems<<<B, 128, 0, s>>>(B, H, nj, dt, Q, stage_w, term_w, theta0, thetag, x0, xref, ff, caug);
  return cudaGetLastError();
}

// RRT route -> CFS reference (RRTstar_CFS.m:96-100):
//   wpTimes = (0:W-1)*dt; trajTimes = linspace(0, wpTimes(end), H+1); sampled = cubicpolytraj(route, wpTimes, trajTimes)
// with the toolbox defaults (zero velocity at every waypoint): on segment k, q = q_k + (3 s^2 - 2 s^3)(q_{k+1} - q_k),
// s = (t - t_k)/(t_{k+1} - t_k).  One CTA per route; outputs theta0 = sampled(:,1), thetag = sampled(:,H+1) and
// x_ = [sampled(:,i); 0] for i = 2..H+1 (RRTstar_CFS.m:106-119).
__global__ void __launch_bounds__(128) k_resample_routes(int B, int Wmax, int H, int nj, double dt, const double *routes /*nj x W x B*/,
                                                         const int *route_len, double *theta0, double *thetag, double *xref) {
  const int b = blockIdx.x, ns = 2 * nj;
  const double *wp = routes + (size_t)b * nj * Wmax;
  const int W = route_len ? (route_len[b] < Wmax ? route_len[b] : Wmax) : Wmax;
  const double t_end = (W - 1) * dt;
  for (int idx = threadIdx.x; idx < (H + 1) * nj; idx += blockDim.x) {
    const int i = idx / nj, k = idx - i * nj;
    if (W < 2) {  // no route (failed RRT seed): a reference that passes the first stop test untouched
      if (i == 0) theta0[(size_t)b * nj + k] = 0.0;
      if (i == H) thetag[(size_t)b * nj + k] = 0.0;
      if (i > 0) {
        xref[(size_t)b * H * ns + (size_t)(i - 1) * ns + k] = 1.0;
        xref[(size_t)b * H * ns + (size_t)(i - 1) * ns + nj + k] = 1.0;
      }
      continue;
    }
    // MATLAB linspace: d1 + (0:n1)*(d2-d1)/n1 with the last point set to d2 exactly
    const double t = (i == H) ? t_end : (i * t_end) / H;
    int seg = 0;  // last waypoint time <= t, clipped to a valid segment
    while (seg < W - 2 && (seg + 1) * dt <= t) ++seg;
    const double t0 = seg * dt, t1 = (seg + 1) * dt;
    const double sn = (t - t0) / (t1 - t0);
    const double blend = 3.0 * sn * sn - 2.0 * sn * sn * sn;
    const double q0 = wp[k + nj * seg], q1 = wp[k + nj * (seg + 1)];
    const double q = q0 + blend * (q1 - q0);
    if (i == 0) theta0[(size_t)b * nj + k] = q;
    if (i == H) thetag[(size_t)b * nj + k] = q;
    if (i > 0) {
      xref[(size_t)b * H * ns + (size_t)(i - 1) * ns + k] = q;
      xref[(size_t)b * H * ns + (size_t)(i - 1) * ns + nj + k] = 0.0;
    }
  }
}

cudaError_t launch_resample_routes(int B, int W, int H, int nj, double dt, const double *routes, const int *route_len,
                                   double *theta0, double *thetag, double *xref, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  k_resample_routes<<<B, 128, 0, s>>>(B, W, H, nj, dt, routes, route_len, theta0, thetag, xref);
  return cudaGetLastError();
}

__global__ void k_mark_no_route(int B, const int *route_len, int *status, int *iters) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && route_len[b] < 2) {
    status[b] = 5;  // CFS_STATUS_NO_ROUTE
    iters[b] = 0;
  }
}
cudaError_t launch_mark_no_route(int B, const int *route_len, int *status, int *iters, cudaStream_t s) {
  if (B <= 0 || !route_len) return cudaSuccess;
  k_mark_no_route<<<(B + 127) / 128, 128, 0, s>>>(B, route_len, status, iters);
  return cudaGetLastError();
}

__global__ void k_route_len_or_fail(int S, const int *route_len, const int *fail, int *out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < S) out[b] = (fail[b] || route_len[b] < 0) ? 0 : route_len[b];
}
cudaError_t launch_route_len_or_fail(int S, const int *route_len, const int *fail, int *out, cudaStream_t s) {
  if (S <= 0) return cudaSuccess;
  k_route_len_or_fail<<<(S + 127) / 128, 128, 0, s>>>(S, route_len, fail, out);
  return cudaGetLastError();
}

cudaError_t setup_gram(int n, int H, int nj, double dt, const double *QQ, double *work_L, double *work_Y, double *G,
                       double *gdiag, int *info, cudaStream_t s) {
  (void)H;
  const int np = 3 * n;
  cudaError_t e;
  if (QQ) {
    k_copy<<<64, 256, 0, s>>>((size_t)n * n, QQ, work_L);
    k_chol<<<1, 1024, 0, s>>>(n, work_L, info);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  } else {
    e = cudaMemsetAsync(info, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
  }
  k_trsm_primitives<<<(np + 63) / 64, 64, 0, s>>>(n, nj, dt, QQ ? work_L : nullptr, work_Y);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  // G = Yt * Yt'  (np x np, K = n):  A(m,k) = Yt[m + np*k], B(k,nn) = Yt[nn + np*k]
  dim3 grid((np + GT - 1) / GT, (np + GT - 1) / GT);
  k_dgemm<<<grid, 256, 0, s>>>(np, np, n, 1.0, work_Y, 1, np, work_Y, np, 1, G, np);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  k_diag<<<(np + 127) / 128, 128, 0, s>>>(np, G, gdiag);
  return cudaGetLastError();
}

}  // namespace cfs
