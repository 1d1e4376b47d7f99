// k_qp.cu -- K2+K3+K4: constraint assembly, batched strictly-convex QP, roll-out, cost and stop rule (sm_100a).
//
// Replaces, per outer iteration and per problem (reference file:line):
//   Lib/CFS_FANUC.m:119-129   get_con row assembly  (l=-Diff'*Bj(1:njoint,:), s=I-Diff'*Bj*u, velocity rows)
//   Lib/CFS_FANUC.m:85        quadprog(QQ,ff,Ainq,binq,[],[],-MAX_input,MAX_input)
//   Lib/PSGCFS_FANUC.m:120    quadprog(I,-u_,Ainq,binq)                       (projection, identity metric)
//   Lib/CFS_FANUC.m:90-94     roll-out xR(:,i)=A*xR(:,i-1)+B*u_i
//   Lib/EVAL.m:51-73          get_cost, store_result, stop_outer
//
// Design (B200-first, not a translation of quadprog):
//   * persistent CTAs pull problems from a device work queue (no host round trip, no tail of idle SMs);
//   * the QP is solved by a dual active-set method (Goldfarb-Idnani) written entirely in "primitive space":
//     v = P u (3n values: joint displacements B_theta u, joint velocities B_omega u, controls u) lives in shared
//     memory; every constraint reads <= nj entries of v, and every metric inner product the method needs is a small
//     bilinear form over the batch-shared Gram matrix G = P QQ^{-1} P' (k_setup.cu) held in L2.  No n x n
//     per-problem factorisation, no dense Ainq (550 x 250 per problem in the reference) is ever formed;
//   * the only per-problem matrix is the inverse of the working-set Gram matrix (q x q, q = #active rows, typically
//     < 8): in shared memory up to q = QP_QS, spilled to a per-CTA global slab beyond, rank-1 updated on add / drop;
//   * the primal point is never stepped: at the top of every outer step v is re-evaluated from the multipliers,
//     v = v0 - G (C_W' lambda), as one pass over the flat list of active (primitive row, weight) terms with many
//     independent L2 loads in flight (small working sets: all 3n entries straight from G; large ones: the n control
//     entries from G followed by the closed-form B_theta / B_omega sums).  The inner loop only needs scalars;
//   * the optimal cost comes from duality: f(u) = f(u0) + 1/2 sum_w lambda_w * (violation of row w at u0).
#include "cfs_kernels.cuh"

namespace cfs {

#define QP_THREADS 128
#define QP_WARPS (QP_THREADS / 32)
#define QP_DEP_TOL 1e-8
#define QP_QS 64          // working sets up to QP_QS keep their inverse in shared memory
#define QP_SMALL_T 16     // term lists up to this length refresh all 3n primitives directly from G

struct QpView {  // decoded shared-memory layout
  double *v;       // np
  double *ocoef;   // OH*nj   (-g)
  double *orhs;    // OH
  double *onrm;    // OH      sqrt(c QQ^-1 c')
  double *lam;     // n+2
  double *r;       // n+2
  double *g;       // n+2
  double *red;     // 16
  double *tcoef;   // TMAX    coefficient of each active term
  double *twgt;    // TMAX    lambda_owner * coefficient
  double *Msm;     // QP_QS*QP_QS
  double *lim;     // 8: velocity limits
  double *w0;      // 8: initial joint velocities
  int *act;        // n+2
  int *toff;       // n+3     first term of each working-set member
  int *trow;       // TMAX    primitive row of each active term
  int *towner;     // TMAX
  int *ctl;        // 8
  unsigned char *inact;  // m
  double *v0s;     // np      v at the unconstrained minimiser (per problem)
  double *gns;     // 2n      QQ^-1 norms of the omega / control primitive rows (per kernel)
  double *ums;     // n       MAX_input (per kernel)
  double *pscr;    // QP_THREADS  partial sums of the polish residual
};

#define QP_NOFF 23
__host__ __device__ inline size_t qp_smem_layout(int n, int nj, int OH, int m, size_t *off /*[QP_NOFF]*/) {
  size_t o = 0;
  const int np = 3 * n;
  const int tmax = nj * OH + n + 2;
  off[0] = o; o += sizeof(double) * np;
  off[1] = o; o += sizeof(double) * (size_t)OH * nj;
  off[2] = o; o += sizeof(double) * OH;
  off[3] = o; o += sizeof(double) * OH;
  off[4] = o; o += sizeof(double) * (n + 2);
  off[5] = o; o += sizeof(double) * (n + 2);
  off[6] = o; o += sizeof(double) * (n + 2);
  off[7] = o; o += sizeof(double) * 16;
  off[8] = o; o += sizeof(double) * tmax;
  off[9] = o; o += sizeof(double) * tmax;
  off[10] = o; o += sizeof(double) * QP_QS * QP_QS;
  off[11] = o; o += sizeof(double) * 8;
  off[12] = o; o += sizeof(double) * 8;
  off[13] = o; o += sizeof(int) * (n + 2);
  off[14] = o; o += sizeof(int) * (n + 4);
  off[15] = o; o += sizeof(int) * tmax;
  off[16] = o; o += sizeof(int) * tmax;
  off[17] = o; o += sizeof(int) * 8;
  off[18] = o; o += (size_t)((m + 15) / 16) * 16;
  off[19] = o; o += sizeof(double) * np;
  off[20] = o; o += sizeof(double) * 2 * n;
  off[21] = o; o += sizeof(double) * n;
  off[22] = o; o += sizeof(double) * QP_THREADS;
  return (o + 15) / 16 * 16;
}

size_t qp_smem_bytes(const SolveArgs &a) {
  size_t off[QP_NOFF];
  const int OH = a.nobs * a.H;
  return qp_smem_layout(a.n, a.nj, OH, OH + 4 * a.n, off);
}

// ---- constraint descriptors --------------------------------------------------------------------------------
// cid in [0,OH): obstacle row (j,i), cid = j*H+i, terms k<nj on theta primitive (i,k) with coefficient ocoef[cid*nj+k]
// cid in [OH,OH+2n): velocity row of omega primitive idx=(cid-OH)>>1, sign bit (0: +row <= lim-w0, 1: -row <= lim+w0)
// cid in [OH+2n,OH+4n): bound row of control idx, sign bit likewise (CFS_FANUC.m:85 lb/ub)
struct Desc {
  int nterm;
  int row0;     // first primitive row; terms are consecutive rows
  double coef;  // single-term coefficient (+-1) when nterm == 1
  const double *cv;  // coefficient vector when nterm > 1
};

__device__ __forceinline__ Desc decode(int cid, int OH, int H, int n, int nj, const double *ocoef) {
  Desc d;
  if (cid < OH) {
    const int i = cid % H;
    d.nterm = nj;
    d.row0 = i * nj;
    d.coef = 0.0;
    d.cv = ocoef + (size_t)cid * nj;
  } else {
    const int e = cid - OH;  // [0,2n): omega primitives n.., [2n,4n): control primitives 2n..
    d.nterm = 1;
    d.row0 = n + (e >> 1);
    d.coef = (e & 1) ? -1.0 : 1.0;
    d.cv = nullptr;
  }
  return d;
}

__device__ __forceinline__ double gram(const Desc &a, const Desc &b, const double *__restrict__ G, int np) {
  if (a.nterm == 1 && b.nterm == 1) return a.coef * b.coef * G[(size_t)a.row0 * np + b.row0];
  double s = 0.0;
  for (int k = 0; k < a.nterm; ++k) {
    const double ca = a.cv ? a.cv[k] : a.coef;
    const double *Gr = G + (size_t)(a.row0 + k) * np + b.row0;
    double t = 0.0;
    for (int l = 0; l < b.nterm; ++l) t += (b.cv ? b.cv[l] : b.coef) * Gr[l];
    s += ca * t;
  }
  return s;
}

// slack = rhs - c u, evaluated from the primitive values v
__device__ __forceinline__ double slack_of(int cid, int OH, int H, int n, int nj, const QpView &s, const double *umax) {
  if (cid < OH) {
    const int i = cid % H;
    const double *c = s.ocoef + (size_t)cid * nj;
    const double *vv = s.v + i * nj;
    double val = 0.0;
    for (int k = 0; k < nj; ++k) val += c[k] * vv[k];
    return s.orhs[cid] - val;
  }
  const int e = cid - OH;
  const int idx = e >> 1, neg = e & 1;
  if (e < 2 * n) {  // CFS_FANUC.m:126-129 : +-Baug_w u <= lim -+ Aaug_w x0
    const int k = idx % nj;
    const double vv = s.v[n + idx];
    return neg ? (s.lim[k] + s.w0[k]) + vv : (s.lim[k] - s.w0[k]) - vv;
  }
  const int c = idx - n;
  const double vv = s.v[2 * n + c];
  return neg ? umax[c] + vv : umax[c] - vv;
}

__device__ __forceinline__ double rhs_scale(int cid, int OH, int n, int nj, const QpView &s, const double *umax) {
  if (cid < OH) return fabs(s.orhs[cid]);
  const int e = cid - OH, idx = e >> 1;
  if (e < 2 * n) return s.lim[idx % nj];
  return umax[idx - n];
}

// c_w u0 - rhs_w : violation of row cid at the unconstrained minimiser (v0s), the right-hand side of S_W lambda = b
__device__ __forceinline__ double viol_at_u0(int cid, int OH, int H, int n, int nj, const QpView &s, const double *umax) {
  if (cid < OH) {
    const int i = cid % H;
    double val = 0.0;
    for (int k = 0; k < nj; ++k) val += s.ocoef[cid * nj + k] * s.v0s[i * nj + k];
    return val - s.orhs[cid];
  }
  const int e = cid - OH, idx = e >> 1, neg = e & 1;
  if (e < 2 * n) {
    const int k = idx % nj;
    return neg ? -s.v0s[n + idx] - (s.lim[k] + s.w0[k]) : s.v0s[n + idx] - (s.lim[k] - s.w0[k]);
  }
  const int c = idx - n;
  return neg ? -s.v0s[2 * n + c] - umax[c] : s.v0s[2 * n + c] - umax[c];
}

// ---- block reductions (QP_THREADS threads) --------------------------------------------------------------------
__device__ __forceinline__ void block_argmin(double &val, int &idx, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, val, o);
    const int oi = __shfl_down_sync(0xffffffffu, idx, o);
    if (ov < val || (ov == val && oi >= 0 && (idx < 0 || oi < idx))) {
      val = ov;
      idx = oi;
    }
  }
  const int w = threadIdx.x >> 5;
  __syncthreads();  // protect red from the previous use
  if ((threadIdx.x & 31) == 0) {
    red[2 * w] = val;
    reinterpret_cast<int *>(red + 2 * w + 1)[0] = idx;
  }
  __syncthreads();
  val = red[0];
  idx = reinterpret_cast<int *>(red + 1)[0];
#pragma unroll
  for (int ww = 1; ww < QP_WARPS; ++ww) {
    const double ov = red[2 * ww];
    const int oi = reinterpret_cast<int *>(red + 2 * ww + 1)[0];
    if (ov < val || (ov == val && oi >= 0 && (idx < 0 || oi < idx))) {
      val = ov;
      idx = oi;
    }
  }
}

__device__ __forceinline__ double block_sum(double val, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = val;
  __syncthreads();
  double s = red[0];
#pragma unroll
  for (int ww = 1; ww < QP_WARPS; ++ww) s += red[ww];
  return s;
}

// ============================================================================================================
// The kernel
// ============================================================================================================
__global__ void __launch_bounds__(QP_THREADS, 3) k_qp(SolveArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, nj = a.nj, H = a.H, np = 3 * n, OH = a.nobs * H, m = OH + 4 * n;
  const int tid = threadIdx.x;
  QpView s;
  {
    size_t off[QP_NOFF];
    qp_smem_layout(n, nj, OH, m, off);
    s.v = reinterpret_cast<double *>(smem_raw + off[0]);
    s.ocoef = reinterpret_cast<double *>(smem_raw + off[1]);
    s.orhs = reinterpret_cast<double *>(smem_raw + off[2]);
    s.onrm = reinterpret_cast<double *>(smem_raw + off[3]);
    s.lam = reinterpret_cast<double *>(smem_raw + off[4]);
    s.r = reinterpret_cast<double *>(smem_raw + off[5]);
    s.g = reinterpret_cast<double *>(smem_raw + off[6]);
    s.red = reinterpret_cast<double *>(smem_raw + off[7]);
    s.tcoef = reinterpret_cast<double *>(smem_raw + off[8]);
    s.twgt = reinterpret_cast<double *>(smem_raw + off[9]);
    s.Msm = reinterpret_cast<double *>(smem_raw + off[10]);
    s.lim = reinterpret_cast<double *>(smem_raw + off[11]);
    s.w0 = reinterpret_cast<double *>(smem_raw + off[12]);
    s.act = reinterpret_cast<int *>(smem_raw + off[13]);
    s.toff = reinterpret_cast<int *>(smem_raw + off[14]);
    s.trow = reinterpret_cast<int *>(smem_raw + off[15]);
    s.towner = reinterpret_cast<int *>(smem_raw + off[16]);
    s.ctl = reinterpret_cast<int *>(smem_raw + off[17]);
    s.inact = smem_raw + off[18];
    s.v0s = reinterpret_cast<double *>(smem_raw + off[19]);
    s.gns = reinterpret_cast<double *>(smem_raw + off[20]);
    s.ums = reinterpret_cast<double *>(smem_raw + off[21]);
    s.pscr = reinterpret_cast<double *>(smem_raw + off[22]);
  }
  const double *__restrict__ G = a.G;
  const double *__restrict__ gnorm = a.gdiag;  // sqrt(diag(G))
  const int has_vel = a.has_lim, has_bnd = a.has_bounds;
  const bool psg = a.solver == 1;
  for (int e = tid; e < 2 * n; e += QP_THREADS) s.gns[e] = gnorm[n + e];
  for (int e = tid; e < n; e += QP_THREADS) s.ums[e] = has_bnd ? a.max_input[e] : 0.0;
  const double *umax = s.ums;
  const double dt = a.tab->dt;
  const int ldg = a.slab_ld;
  double *Mgl = a.slab + (size_t)blockIdx.x * ldg * ldg;  // spill area of the working-set inverse
  const int count = *a.count_cur;
  long long steps_total = 0;
  int qmax_seen = 0;
  // optional phase profile (thread 0 clocks): 0 prologue, 1 refresh, 2 scan, 3 gram+solve+steplen, 4 update, 5 epilogue,
  // 6 problems, 7 outer steps
  long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tck = 0;
#define PF_START() do { if (a.prof && tid == 0) tck = clock64(); } while (0)
#define PF_ADD(k) do { if (a.prof && tid == 0) { const long long now_ = clock64(); pf[k] += now_ - tck; tck = now_; } } while (0)

  for (;;) {
    __syncthreads();
    if (tid == 0) s.ctl[0] = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int slot = s.ctl[0];
    if (slot >= count) break;
    const int b = a.list_cur[slot];
    const double *x0 = a.x0 + (size_t)b * 2 * nj;
    double *ub = a.u + (size_t)b * n;
    double *xb = a.x + (size_t)b * 2 * n;
    const double *v0 = a.v0 + (size_t)b * np;

    // ---- prologue: v = v0; disp = B_theta u_cur; obstacle rows ------------------------------------------------
    PF_START();
    for (int pi = tid; pi < np; pi += QP_THREADS) {
      const double t_ = v0[pi];
      s.v0s[pi] = t_;
      s.v[pi] = t_;
    }
    for (int e = tid; e < m; e += QP_THREADS) s.inact[e] = 0;
    if (tid < 8) {
      s.lim[tid] = (tid < nj && has_vel) ? a.lim[tid] : 0.0;
      s.w0[tid] = (tid < nj) ? x0[nj + tid] : 0.0;
    }
    // disp(i,k) = (Bj(1:njoint,:)*u)(k) of CFS_FANUC.m:120 -> s.g.  In the first outer iteration u = 0 (CFS_FANUC.m:56)
    // while x_ is the reference line; afterwards x_ is the roll-out of u, so B_theta u = theta_i - (theta_0 + i dt w_0).
    for (int e = tid; e < n; e += QP_THREADS) {
      const int i = e / nj, k = e % nj;
      s.g[e] = (a.outer_iter == 1) ? 0.0 : xb[(size_t)i * 2 * nj + k] - (x0[k] + ((i + 1) * dt) * x0[nj + k]);
    }
    __syncthreads();
    for (int cid = tid; cid < OH; cid += QP_THREADS) {
      const int j = cid / H, i = cid % H;
      const double *gr = a.grad + ((size_t)b * OH + cid) * nj;
      const double dist = a.dist[(size_t)b * OH + cid];
      const double margin = a.margin_is_D ? a.tab->obs[j].D : a.tab->obs[j].eps;
      double gu = 0.0;
      for (int k = 0; k < nj; ++k) {
        s.ocoef[cid * nj + k] = -gr[k];                 // l = -Diff'*Bj(1:njoint,:)   (CFS_FANUC.m:121)
        gu += gr[k] * s.g[i * nj + k];
      }
      s.orhs[cid] = (dist - margin) - gu;               // s = I - Diff'*Bj*u          (CFS_FANUC.m:117,120)
    }
    __syncthreads();
    for (int cid = tid; cid < OH; cid += QP_THREADS) {
      const Desc d = decode(cid, OH, H, n, nj, s.ocoef);
      const double sg = gram(d, d, G, np);
      s.onrm[cid] = sg > 0.0 ? sqrt(sg) : 0.0;
    }
    if (tid == 0) s.toff[0] = 0;
    __syncthreads();
    PF_ADD(0);
    pf[6] += 1;

    // ---- dual active-set iterations --------------------------------------------------------------------------
    int q = 0, status = -1, steps = 0;
    bool in_smem = true, polished = false;
    if (psg && a.skip[b]) {  // stop_inner() already true: u stays, no projection (PSGCFS_FANUC.m:88)
      for (int c = tid; c < n; c += QP_THREADS) s.v[2 * n + c] = ub[c];
      __syncthreads();
      status = 0;
    }
    double fval = a.cost0[b];
    const double fupper = (has_bnd && a.fupper) ? a.fupper[b] : INFINITY;
    const int max_steps = 20 * (m + n) + 100;
#define MAT(r_, c_) (in_smem ? s.Msm[(r_) + QP_QS * (c_)] : Mgl[(r_) + (size_t)ldg * (c_)])
    while (status < 0) {
      // (0) primal recovery from the multipliers: v = v0 - G (C_W' lambda)
      if (q > 0) {
        const int T = s.toff[q];
        for (int t = tid; t < T; t += QP_THREADS) s.twgt[t] = s.lam[s.towner[t]] * s.tcoef[t];
        __syncthreads();
        if (T <= QP_SMALL_T) {
          for (int base = 0; base < np; base += 6 * QP_THREADS) {  // 6 primitives per thread, 2 terms per pass: 12 loads in flight
            double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            int t = 0;
            for (; t + 2 <= T; t += 2) {
              const double w0_ = s.twgt[t], w1_ = s.twgt[t + 1];
              const double *g0 = G + (size_t)s.trow[t] * np + base + tid, *g1 = G + (size_t)s.trow[t + 1] * np + base + tid;
              double l0[6], l1[6];
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                const bool ok = base + tid + j * QP_THREADS < np;
                l0[j] = ok ? g0[j * QP_THREADS] : 0.0;
                l1[j] = ok ? g1[j * QP_THREADS] : 0.0;
              }
#pragma unroll
              for (int j = 0; j < 6; ++j) acc[j] += w0_ * l0[j] + w1_ * l1[j];
            }
            if (t < T) {
              const double w0_ = s.twgt[t];
              const double *g0 = G + (size_t)s.trow[t] * np + base + tid;
#pragma unroll
              for (int j = 0; j < 6; ++j)
                if (base + tid + j * QP_THREADS < np) acc[j] += w0_ * g0[j * QP_THREADS];
            }
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const int pi = base + tid + j * QP_THREADS;
              if (pi < np) s.v[pi] = s.v0s[pi] - acc[j];
            }
          }
        } else {
          const double *__restrict__ Gu = G + 2 * n;  // control block of every primitive row
          for (int c = tid; c < n; c += QP_THREADS) {
            double acc8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int t = 0;
            for (; t + 8 <= T; t += 8) {
              double ld8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) ld8[j] = Gu[(size_t)s.trow[t + j] * np + c];
#pragma unroll
              for (int j = 0; j < 8; ++j) acc8[j] += s.twgt[t + j] * ld8[j];
            }
            for (; t < T; ++t) acc8[0] += s.twgt[t] * Gu[(size_t)s.trow[t] * np + c];
            s.v[2 * n + c] = s.v0s[2 * n + c] - (((acc8[0] + acc8[1]) + (acc8[2] + acc8[3])) + ((acc8[4] + acc8[5]) + (acc8[6] + acc8[7])));
          }
          __syncthreads();
          for (int e = tid; e < n; e += QP_THREADS) {  // B_theta u and B_omega u in closed form
            const int i = e / nj, k = e % nj;
            double at = 0.0, aw = 0.0;
            for (int j = 0; j <= i; ++j) {
              const double uj = s.v[2 * n + j * nj + k];
              at += (0.5 * dt * dt + ((i - j) * dt) * dt) * uj;
              aw += dt * uj;
            }
            s.v[e] = at;
            s.v[n + e] = aw;
          }
        }
        __syncthreads();
      }
      PF_ADD(1);
      pf[7] += 1;
      // (1) most violated inactive row, normalised by its QQ^-1 norm
      double best = 0.0;
      int bidx = -1;
      for (int cid = tid; cid < OH; cid += QP_THREADS) {
        const double nr = s.onrm[cid];
        if (s.inact[cid] || !(nr > 0.0)) continue;
        const double sl = slack_of(cid, OH, H, n, nj, s, umax);
        if (sl < -1e-11 * (1.0 + fabs(s.orhs[cid]))) {
          const double val = sl / nr;
          if (val < best || bidx < 0) {
            best = val;
            bidx = cid;
          }
        }
      }
      // omega / control primitives: both signs of a row share its value and its norm
      for (int e = tid; e < 2 * n; e += QP_THREADS) {
        const bool is_w = e < n;
        if (is_w ? !has_vel : !has_bnd) continue;
        const double nr = s.gns[e];
        if (!(nr > 0.0)) continue;
        const double vv = s.v[n + e];
        double up, lo, sc;
        if (is_w) {
          const int k = e % nj;
          up = (s.lim[k] - s.w0[k]) - vv;
          lo = (s.lim[k] + s.w0[k]) + vv;
          sc = s.lim[k];
        } else {
          up = umax[e - n] - vv;
          lo = umax[e - n] + vv;
          sc = umax[e - n];
        }
        const double tol = 1e-11 * (1.0 + sc);
        const int cu = OH + 2 * e;
        if (up < -tol && !s.inact[cu]) {
          const double val = up / nr;
          if (val < best || bidx < 0) {
            best = val;
            bidx = cu;
          }
        }
        if (lo < -tol && !s.inact[cu + 1]) {
          const double val = lo / nr;
          if (val < best || bidx < 0) {
            best = val;
            bidx = cu + 1;
          }
        }
      }
      block_argmin(best, bidx, s.red);
      PF_ADD(2);
      if (bidx < 0) {
        if (q == 0 || polished) {
          status = 0;
          break;
        }
        // Polish: the working-set inverse M has been rank-1 updated `steps` times; one step of iterative refinement on
        // S_W lambda = b (S_W = C_W QQ^-1 C_W' re-read from G, b = violations at u0) removes the accumulated drift
        // (measured: 7e-10 -> 3e-13 in u).  Then v is re-evaluated from the refined multipliers and scanned once more.
        {
          const int nch = QP_THREADS / q > 0 ? QP_THREADS / q : 1;  // chunks of columns per row, fixed summation order
          if (q <= QP_THREADS) {
            const int w = tid % q, ch = tid / q;
            if (ch < nch) {
              const Desc dw = decode(s.act[w], OH, H, n, nj, s.ocoef);
              double acc = 0.0;
              for (int c = ch; c < q; c += nch) acc += gram(dw, decode(s.act[c], OH, H, n, nj, s.ocoef), G, np) * s.lam[c];
              s.pscr[ch * q + w] = acc;
            }
            __syncthreads();
            if (tid < q) {
              double acc = 0.0;
              for (int ch2 = 0; ch2 < nch; ++ch2) acc += s.pscr[ch2 * q + tid];
              s.g[tid] = viol_at_u0(s.act[tid], OH, H, n, nj, s, umax) - acc;
            }
          } else {
            for (int w = tid; w < q; w += QP_THREADS) {
              const Desc dw = decode(s.act[w], OH, H, n, nj, s.ocoef);
              double acc = 0.0;
              for (int c = 0; c < q; ++c) acc += gram(dw, decode(s.act[c], OH, H, n, nj, s.ocoef), G, np) * s.lam[c];
              s.g[w] = viol_at_u0(s.act[w], OH, H, n, nj, s, umax) - acc;
            }
          }
          __syncthreads();
          for (int w = tid; w < q; w += QP_THREADS) {
            double acc = 0.0;
            for (int c = 0; c < q; ++c) acc += MAT(w, c) * s.g[c];
            s.r[w] = acc;
          }
          __syncthreads();
          for (int w = tid; w < q; w += QP_THREADS) s.lam[w] += s.r[w];
          __syncthreads();
        }
        polished = true;
        continue;
      }
      polished = false;
      const int p = bidx;
      const Desc dp = decode(p, OH, H, n, nj, s.ocoef);
      const double sigma = gram(dp, dp, G, np);
      double sp = slack_of(p, OH, H, n, nj, s, umax);
      double lam_p = 0.0;
      // (2) bring row p into the working set
      for (;;) {
        if (++steps > max_steps) {
          status = 3;
          break;
        }
        // g_w = c_w QQ^-1 c_p'
        for (int w = tid; w < q; w += QP_THREADS) s.g[w] = gram(decode(s.act[w], OH, H, n, nj, s.ocoef), dp, G, np);
        __syncthreads();
        // r = Minv g
        double part = 0.0;
        for (int w = tid; w < q; w += QP_THREADS) {
          double acc8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          int c = 0;
          for (; c + 8 <= q; c += 8) {
            double ld8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ld8[j] = MAT(w, c + j);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc8[j] += ld8[j] * s.g[c + j];
          }
          for (; c < q; ++c) acc8[0] += MAT(w, c) * s.g[c];
          const double acc = ((acc8[0] + acc8[1]) + (acc8[2] + acc8[3])) + ((acc8[4] + acc8[5]) + (acc8[6] + acc8[7]));
          s.r[w] = acc;
          part += s.g[w] * acc;
        }
        const double delta = sigma - block_sum(part, s.red);  // z'n+ in Goldfarb-Idnani's notation
        // t1: largest dual step keeping the multipliers non-negative
        double t1 = INFINITY;
        int l = -1;
        for (int w = tid; w < q; w += QP_THREADS)
          if (s.r[w] > 0.0) {
            const double t = s.lam[w] / s.r[w];
            if (t < t1 || l < 0) {
              t1 = t;
              l = w;
            }
          }
        block_argmin(t1, l, s.red);
        if (l < 0) t1 = INFINITY;
        if (!(delta == delta) || !(sigma == sigma)) {
          status = 3;
          break;
        }
        // Row p is treated as linearly dependent on the working set when its QQ^-1-orthogonal remainder is below
        // 1e-8 of its norm^2: in Gram form delta carries cancellation noise ~eps*cond(S_W)*sigma, and a step of
        // length -sp/delta along such a direction only manufactures astronomically large multipliers.
        const bool dependent = !(delta > QP_DEP_TOL * sigma) || q >= n;
        double t2 = INFINITY;
        if (!dependent) {
          t2 = -sp / delta;
          if (t2 < 0.0) t2 = 0.0;
        }
        if (l < 0 && dependent) {
          status = 2;  // infeasible
          break;
        }
        const bool full = (t2 <= t1);
        const double t = full ? t2 : t1;
        // dual objective (Goldfarb-Idnani: f += t z'n+ (t/2 + u+_{q+1})); weak duality: if it exceeds an upper bound of
        // the primal objective over the box |u| <= MAX_input the QP has no feasible point.
        if (!dependent) {
          fval += t * delta * (0.5 * t + lam_p);
          sp += t * delta;  // slack of p moves by t z'n+
        }
        if (fval > fupper) {
          status = 2;
          break;
        }
        for (int w = tid; w < q; w += QP_THREADS) s.lam[w] -= t * s.r[w];
        lam_p += t;
        __syncthreads();
        PF_ADD(3);
        if (full) {
          if (in_smem && q + 1 > QP_QS) {  // spill the inverse to the global slab
            for (int e = tid; e < q * q; e += QP_THREADS) Mgl[(e % q) + (size_t)ldg * (e / q)] = s.Msm[(e % q) + QP_QS * (e / q)];
            in_smem = false;
            __syncthreads();
          }
          // add p: bordered inverse  [[M + r r'/d, -r/d], [-r'/d, 1/d]]
          const double id = 1.0 / delta;
          {
            const int q1 = q + 1, tot = q1 * q1;
            for (int e0 = tid; e0 < tot; e0 += 4 * QP_THREADS) {
              double old4[4];
              int rr4[4], cc4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = e0 + j * QP_THREADS;
                rr4[j] = e % q1;
                cc4[j] = e / q1;
                old4[j] = (e < tot && rr4[j] < q && cc4[j] < q) ? MAT(rr4[j], cc4[j]) : 0.0;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = e0 + j * QP_THREADS;
                if (e >= tot) continue;
                const int r_ = rr4[j], c_ = cc4[j];
                double val;
                if (r_ < q && c_ < q)
                  val = old4[j] + s.r[r_] * s.r[c_] * id;
                else if (r_ == q && c_ == q)
                  val = id;
                else
                  val = -s.r[r_ < q ? r_ : c_] * id;
                MAT(r_, c_) = val;
              }
            }
          }
          const int t0 = s.toff[q];
          if (tid < dp.nterm) {
            s.trow[t0 + tid] = dp.row0 + tid;
            s.tcoef[t0 + tid] = dp.cv ? dp.cv[tid] : dp.coef;
            s.towner[t0 + tid] = q;
          }
          if (tid == 0) {
            s.act[q] = p;
            s.lam[q] = lam_p;
            s.inact[p] = 1;
            s.toff[q + 1] = t0 + dp.nterm;
          }
          ++q;
          if (q > qmax_seen) qmax_seen = q;
          __syncthreads();
          PF_ADD(4);
          break;
        }
        // drop working-set member l: M <- M - M(:,l) M(l,:)/M(l,l), then move the last member into slot l
        {
          const int last = q - 1;
          for (int w = tid; w < q; w += QP_THREADS) s.g[w] = MAT(w, l);
          __syncthreads();
          const double ip = 1.0 / s.g[l];
          for (int e0 = tid; e0 < q * q; e0 += 4 * QP_THREADS) {
            double old4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = e0 + j * QP_THREADS;
              old4[j] = e < q * q ? MAT(e % q, e / q) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = e0 + j * QP_THREADS;
              if (e < q * q) MAT(e % q, e / q) = old4[j] - s.g[e % q] * s.g[e / q] * ip;
            }
          }
          __syncthreads();
          if (l != last) {
            for (int w = tid; w < q; w += QP_THREADS) MAT(w, l) = MAT(w, last);
            __syncthreads();
            for (int w = tid; w < q; w += QP_THREADS) MAT(l, w) = MAT(last, w);
          }
          if (tid == 0) {
            s.inact[s.act[l]] = 0;
            s.act[l] = s.act[last];
            s.lam[l] = s.lam[last];
            int o = 0;  // rebuild the term offsets (drops are rare)
            for (int w = 0; w < last; ++w) {
              s.toff[w] = o;
              o += (s.act[w] < OH) ? nj : 1;
            }
            s.toff[last] = o;
          }
          --q;
          __syncthreads();
          for (int w = tid; w < q; w += QP_THREADS) {
            const Desc d = decode(s.act[w], OH, H, n, nj, s.ocoef);
            const int t0 = s.toff[w];
            for (int k = 0; k < d.nterm; ++k) {
              s.trow[t0 + k] = d.row0 + k;
              s.tcoef[t0 + k] = d.cv ? d.cv[k] : d.coef;
              s.towner[t0 + k] = w;
            }
          }
          __syncthreads();
          PF_ADD(4);
        }
      }
    }
#undef MAT
    steps_total += steps;
    if (tid == 0 && a.prob_steps) a.prob_steps[b] += steps;

    // ---- epilogue ------------------------------------------------------------------------------------------------
    const int it = a.outer_iter;  // 1-based
    if (status == 0) {
      // u = control block of v; e_u = ||u_old - u||   (EVAL.m:58)
      double pe = 0.0;
      for (int c = tid; c < n; c += QP_THREADS) {
        const double un = s.v[2 * n + c];
        const double dlt = ub[c] - un;
        pe += dlt * dlt;
        ub[c] = un;
      }
      const double e_u = sqrt(block_sum(pe, s.red));
      // cost by duality: f(u) = f(u0) + 1/2 sum_w lam_w * (c_w u0 - rhs_w)
      double pc = 0.0;
      for (int w = tid; w < q; w += QP_THREADS) {
        const double val0 = viol_at_u0(s.act[w], OH, H, n, nj, s, umax);
        pc += s.lam[w] * val0;
      }
      const double cost = psg ? 0.0 : a.cost0[b] + 0.5 * block_sum(pc, s.red);  // PSGCFS: k_psg_cost
      // roll-out (CFS_FANUC.m:90-94): x_old is staged in shared memory (coalesced), nj threads run the recurrence
      // xR(:,i) = A xR(:,i-1) + B u_i there, then the new trajectory is written back coalesced; ||x_new - x_old||^2
      // (EVAL.m:64) is accumulated on the way.  xbuf aliases the term-weight scratch (tcoef|twgt), free by now.
      double *xbuf = s.tcoef;
      __syncthreads();
      for (int e = tid; e < 2 * n; e += QP_THREADS) xbuf[e] = xb[e];
      __syncthreads();
      double px = 0.0;
      if (tid < nj) {
        double th = x0[tid], om = x0[nj + tid];
        for (int i = 0; i < H; ++i) {
          const double uk = s.v[2 * n + i * nj + tid];
          const double thn = (th + dt * om) + (0.5 * dt * dt) * uk;
          const double omn = om + dt * uk;
          th = thn;
          om = omn;
          double *xs = xbuf + (size_t)i * 2 * nj;
          // CFS: x_old = previous x_ (CFS_FANUC.m:88); PSGCFS never updates eval.x_old, it stays ones (EVAL.m:47)
          const double d1 = th - (psg ? 1.0 : xs[tid]), d2 = om - (psg ? 1.0 : xs[nj + tid]);
          px += d1 * d1 + d2 * d2;
          xs[tid] = th;
          xs[nj + tid] = om;
        }
      }
      __syncthreads();
      for (int e = tid; e < 2 * n; e += QP_THREADS) xb[e] = xbuf[e];
      const double dx = sqrt(block_sum(px, s.red));
      if (tid == 0) {
        if (!psg) a.cost_hist[(size_t)b * a.max_outer + (it - 1)] = cost;
        if (a.e_u_hist) a.e_u_hist[(size_t)b * a.max_outer + (it - 1)] = e_u;
        a.iters[b] = it;
        if (dx < a.eps_outer) {
          a.status[b] = 0;  // converged (EVAL.m:64-67)
        } else if (it + 1 > a.max_outer) {
          a.status[b] = 1;  // MAX_ITER (EVAL.m:69-72)
        } else {
          a.list_next[atomicAdd(a.count_next, 1)] = b;
        }
      }
    } else if (tid == 0) {
      a.status[b] = status;  // 2 infeasible / 3 numerical: u, x keep the previous iterate
    }
    PF_ADD(5);
  }
  if (a.prof && tid == 0)
    for (int k = 0; k < 8; ++k)
      if (pf[k]) atomicAdd(reinterpret_cast<unsigned long long *>(a.prof + k), (unsigned long long)pf[k]);
  if (tid == 0) {
    if (steps_total) atomicAdd(reinterpret_cast<unsigned long long *>(a.qp_steps), (unsigned long long)steps_total);
    if (qmax_seen) atomicMax(a.max_active, qmax_seen);
  }
}

int qp_max_grid(const SolveArgs &a, int device) {
  int sms = 0, per = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const size_t smem = qp_smem_bytes(a);
  cudaFuncSetAttribute(k_qp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_qp, QP_THREADS, smem);
  if (per < 1) per = 1;
  return sms * per;
}

cudaError_t launch_qp(const SolveArgs &a, int grid, cudaStream_t st) {
  const size_t smem = qp_smem_bytes(a);
  k_qp<<<grid, QP_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

// ============================================================================================================
// Batch initialisation
// ============================================================================================================
// one CTA per problem: u=0, x_=sys_info.x_ (CFS_FANUC.m:55-56), histories NaN, first stop_outer test with
// x_old = ones (EVAL.m:47,61-73), initial active list.
__global__ void __launch_bounds__(128) k_solve_init(SolveArgs a) {
  __shared__ double red[QP_WARPS];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, N = 2 * n;
  double part = 0.0;
  for (int e = tid; e < N; e += blockDim.x) {
    const double xv = a.xref[(size_t)b * N + e];
    a.x[(size_t)b * N + e] = xv;
    part += (xv - 1.0) * (xv - 1.0);
  }
  for (int e = tid; e < n; e += blockDim.x) a.u[(size_t)b * n + e] = 0.0;
  for (int e = tid; e < a.max_outer; e += blockDim.x) {
    a.cost_hist[(size_t)b * a.max_outer + e] = __longlong_as_double(0x7ff8000000000000LL);
    if (a.e_u_hist) a.e_u_hist[(size_t)b * a.max_outer + e] = __longlong_as_double(0x7ff8000000000000LL);
  }
  const double nrm = sqrt(block_sum(part, red));
  if (tid == 0) {
    a.iters[b] = 0;
    a.flags[b] = 0;
    if (a.solver == 1) {
      a.cost_old[b] = 100000.0;   // EVAL.m:29
      a.cost_new[b] = a.caug[b];  // get_cost(u = 0), PSGCFS_FANUC.m:66
      a.skip[b] = 0;
    }
    if (nrm < a.eps_outer) {
      a.status[b] = 0;
    } else if (1 > a.max_outer) {
      a.status[b] = 1;
    } else {
      a.status[b] = -1;
      a.list_cur[atomicAdd(a.count_cur, 1)] = b;
    }
  }
}

cudaError_t launch_solve_init(const SolveArgs &a, cudaStream_t s) {
  if (a.B <= 0) return cudaSuccess;
  k_solve_init<<<a.B, 128, 0, s>>>(a);
  return cudaGetLastError();
}

// v0 = P u0 and cost0 = caug + 1/2 ff'u0 (value of EVAL.get_cost at the unconstrained minimiser), one CTA per problem
__global__ void __launch_bounds__(128) k_v0(SolveArgs a) {
  __shared__ double red[QP_WARPS];
  extern __shared__ double u0s[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, nj = a.nj, np = 3 * n;
  const double dt = a.tab->dt;
  double part = 0.0;
  for (int e = tid; e < n; e += blockDim.x) {
    const double uv = a.u0[(size_t)b * n + e];
    u0s[e] = uv;
    part += a.ff[(size_t)b * n + e] * uv;
  }
  const double fu = block_sum(part, red);
  if (tid == 0) a.cost0[b] = a.caug[b] + 0.5 * fu;
  if (a.fupper) {  // sup over the box |u|<=MAX_input of EVAL.get_cost: 1/2 ||QQ||_inf sum umax^2 + sum |ff| umax + caug
    double pb = 0.0;
    if (a.has_bounds)
      for (int e = tid; e < n; e += blockDim.x) {
        const double um = a.max_input[e];
        pb += 0.5 * a.qq_norm_inf * um * um + fabs(a.ff[(size_t)b * n + e]) * um;
      }
    const double sb = block_sum(pb, red);
    if (tid == 0) a.fupper[b] = a.has_bounds ? (fabs(a.caug[b]) + sb) * (1.0 + 1e-6) + 1e-6 : INFINITY;
  }
  for (int pi = tid; pi < np; pi += blockDim.x) {
    double acc;
    if (pi < n) {
      const int i = pi / nj, k = pi % nj;
      acc = 0.0;
      for (int j = 0; j <= i; ++j) acc += (0.5 * dt * dt + ((i - j) * dt) * dt) * u0s[j * nj + k];
    } else if (pi < 2 * n) {
      const int qq = pi - n, i = qq / nj, k = qq % nj;
      acc = 0.0;
      for (int j = 0; j <= i; ++j) acc += dt * u0s[j * nj + k];
    } else {
      acc = u0s[pi - 2 * n];
    }
    a.v0[(size_t)b * np + pi] = acc;
  }
}

cudaError_t launch_v0(const SolveArgs &a, cudaStream_t s) {
  if (a.B <= 0) return cudaSuccess;
  k_v0<<<a.B, 128, sizeof(double) * a.n, s>>>(a);
  return cudaGetLastError();
}

// problems still marked running (-1) after the last launched iteration cannot exist; flags are merged into status here
__global__ void k_finalize(SolveArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  int st = a.status[b];
  if (st < 0) st = 1;
  a.status[b] = st | (a.flags[b] & 0x100);
}

cudaError_t launch_finalize(const SolveArgs &a, cudaStream_t s) {
  if (a.B <= 0) return cudaSuccess;
  k_finalize<<<(a.B + 127) / 128, 128, 0, s>>>(a);
  return cudaGetLastError();
}

// ---- K5: PSG_update_arm (PSGCFS_FANUC.m:106-112), one CTA per active problem ------------------------------------
//   u_ = u - alpha*((QQ*u + ff) + 10*noise/(iter_O^2+1));   then v0 = P u_ for the projection QP (Hessian I, no bounds).
// stop_inner (PSGCFS_FANUC.m:136-142) is evaluated first: |cost_new - cost_old| < epsilon_I = 1e-4 skips the step.
__global__ void __launch_bounds__(128) k_psg_point(SolveArgs a) {
  extern __shared__ double u0s[];
  if ((int)blockIdx.x >= *a.count_cur) return;
  const int b = a.list_cur[blockIdx.x], tid = threadIdx.x;
  const int n = a.n, nj = a.nj, np = 3 * n;
  const double dt = a.tab->dt;
  const double cn = a.cost_new[b], co = a.cost_old[b];
  if (fabs(cn - co) < 1e-4) {  // uniform across the CTA
    if (tid == 0) a.skip[b] = 1;
    return;
  }
  const double it = (double)a.outer_iter;
  const double *nz = a.noise ? a.noise + ((size_t)b * a.max_outer + (a.outer_iter - 1)) * n : nullptr;
  for (int e = tid; e < n; e += blockDim.x) {
    const double g = a.w[(size_t)b * n + e] + a.ff[(size_t)b * n + e];
    const double uv = a.u[(size_t)b * n + e] - a.alpha * (g + 10 * (nz ? nz[e] : 0.0) / (it * it + 1));
    u0s[e] = uv;
    a.u0[(size_t)b * n + e] = uv;
  }
  if (tid == 0) {
    a.skip[b] = 0;
    a.cost_old[b] = cn;  // PSGCFS_FANUC.m:89
    a.cost0[b] = 0.0;
  }
  __syncthreads();
  for (int pi = tid; pi < np; pi += blockDim.x) {
    double acc;
    if (pi < n) {
      const int i = pi / nj, k = pi % nj;
      acc = 0.0;
      for (int j = 0; j <= i; ++j) acc += (0.5 * dt * dt + ((i - j) * dt) * dt) * u0s[j * nj + k];
    } else if (pi < 2 * n) {
      const int qq = pi - n, i = qq / nj, k = qq % nj;
      acc = 0.0;
      for (int j = 0; j <= i; ++j) acc += dt * u0s[j * nj + k];
    } else {
      acc = u0s[pi - 2 * n];
    }
    a.v0[(size_t)b * np + pi] = acc;
  }
}

cudaError_t launch_psg_point(const SolveArgs &a, cudaStream_t s) {
  if (a.B <= 0) return cudaSuccess;
  k_psg_point<<<a.B, 128, sizeof(double) * a.n, s>>>(a);
  return cudaGetLastError();
}

// EVAL.get_cost (EVAL.m:51-53) of the projected iterate, w = QQ*u from the batched GEMM; one CTA per problem of the list
// the QP kernel just consumed.  Problems whose projection failed this iteration (iters[b] != outer_iter) are skipped.
__global__ void __launch_bounds__(128) k_psg_cost(SolveArgs a) {
  __shared__ double red[QP_WARPS];
  if ((int)blockIdx.x >= *a.count_cur) return;
  const int b = a.list_cur[blockIdx.x], tid = threadIdx.x, n = a.n;
  if (a.iters[b] != a.outer_iter) return;
  double pq = 0.0, pl = 0.0;
  for (int e = tid; e < n; e += blockDim.x) {
    const double uv = a.u[(size_t)b * n + e];
    pq += uv * a.w[(size_t)b * n + e];
    pl += uv * a.ff[(size_t)b * n + e];
  }
  const double quad = block_sum(pq, red);
  const double lin = block_sum(pl, red);
  if (tid == 0) {
    const double cost = (0.5 * quad + lin) + a.caug[b];
    a.cost_new[b] = cost;
    a.cost_hist[(size_t)b * a.max_outer + (a.outer_iter - 1)] = cost;
  }
}

cudaError_t launch_psg_cost(const SolveArgs &a, cudaStream_t s) {
  if (a.B <= 0) return cudaSuccess;
  k_psg_cost<<<a.B, 128, 0, s>>>(a);
  return cudaGetLastError();
}

// ============================================================================================================
// Dense get_con rows for one problem, in the reference's order (CFS_FANUC.m:110-131); parity/debug entry.
// ============================================================================================================
__global__ void k_get_con_rows(const DevTables *tab, int H, int nj, int nobs, int has_lim, int margin_is_D,
                               const double *x0, const double *u, const double *lim, const double *dist,
                               const double *grad, double *Ainq, double *binq, int m) {
  const int n = H * nj;
  const int per = has_lim ? 1 + 2 * nj : 1;
  const double dt = tab->dt;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)m * n) return;
  const int row = (int)(e % m), col = (int)(e / m);
  const int blk = row / per, sub = row % per;
  const int j = blk / H, i = blk % H;
  const int jj = col / nj, k = col % nj;
  double val = 0.0;
  if (sub == 0) {
    const double g = grad[((size_t)j * H + i) * nj + k];
    if (jj <= i) val = -(g * (0.5 * dt * dt + ((i - jj) * dt) * dt));
    if (col == 0) {
      double gu = 0.0;
      for (int c = 0; c < n; ++c) {
        const int j2 = c / nj, k2 = c % nj;
        if (j2 <= i) gu += (grad[((size_t)j * H + i) * nj + k2] * (0.5 * dt * dt + ((i - j2) * dt) * dt)) * u[c];
      }
      const double margin = margin_is_D ? tab->obs[j].D : tab->obs[j].eps;
      binq[row] = (dist[(size_t)j * H + i] - margin) - gu;
    }
  } else {
    const int kk = (sub - 1) % nj, neg = (sub - 1) / nj;
    if (kk == k && jj <= i) val = neg ? -dt : dt;
    if (col == 0) binq[row] = neg ? lim[kk] + x0[nj + kk] : lim[kk] - x0[nj + kk];
  }
  Ainq[row + (size_t)m * col] = val;
}

cudaError_t launch_get_con_rows(const DevTables *tab, int H, int nj, int nobs, int has_lim, int margin_is_D,
                                const double *x0, const double *u, const double *lim, const double *dist,
                                const double *grad, double *Ainq, double *binq, int m, cudaStream_t s) {
  const long long total = (long long)m * H * nj;
  if (total <= 0) return cudaSuccess;
  k_get_con_rows<<<(int)((total + 255) / 256), 256, 0, s>>>(tab, H, nj, nobs, has_lim, margin_is_D, x0, u, lim, dist,
                                                              grad, Ainq, binq, m);
  return cudaGetLastError();
}

}  // namespace cfs
