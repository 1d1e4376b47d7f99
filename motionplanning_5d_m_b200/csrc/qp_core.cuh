// qp_core.cuh -- the batched dual active-set QP core shared by the lock-step kernel (k_qp.cu) and the fused persistent
// kernel (k_fused.cu).  See k_qp.cu for the design notes.
#pragma once
#include "cfs_kernels.cuh"

namespace cfs {

#define QP_THREADS 128   // lock-step kernel / bulk tier of the fused kernel
#define QP_DEP_TOL 1e-8
#define QP_QS 48          // working sets up to QP_QS keep their inverse in shared memory
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
__host__ __device__ inline size_t qp_smem_layout(int n, int nj, int OH, int m, size_t *off /*[QP_NOFF]*/, int qs = QP_QS,
                                                  int nt = QP_THREADS) {
  size_t o = 0;
  const int np = 3 * n;
  const int tmax = nj * OH + n + 2;
  // [Msm | v | tcoef | twgt | lam | r | g] first and contiguous: all of it is dead while the fused kernel runs its
  // gradient phase, which aliases its sin/cos cache onto this span (qp_scratch_span()).
  off[10] = o; o += sizeof(double) * qs * qs;
  off[0] = o; o += sizeof(double) * np;
  off[8] = o; o += sizeof(double) * tmax;
  off[9] = o; o += sizeof(double) * tmax;
  off[4] = o; o += sizeof(double) * (n + 2);
  off[5] = o; o += sizeof(double) * (n + 2);
  off[6] = o; o += sizeof(double) * (n + 2);
  off[1] = o; o += sizeof(double) * (size_t)OH * nj;
  off[2] = o; o += sizeof(double) * OH;
  off[3] = o; o += sizeof(double) * OH;
  off[7] = o; o += sizeof(double) * 64;
  off[11] = o; o += sizeof(double) * 8;
  off[12] = o; o += sizeof(double) * 8;
  off[13] = o; o += sizeof(int) * (n + 2);
  off[14] = o; o += sizeof(int) * (n + 4);
  off[15] = o; o += sizeof(int) * tmax;
  off[16] = o; o += sizeof(int) * tmax;
  off[17] = o; o += sizeof(int) * 8;
  off[18] = o; o += (size_t)((m + 15) / 16) * 16;
  o = (o + 15) / 16 * 16;
  off[19] = o; o += sizeof(double) * np;
  off[20] = o; o += sizeof(double) * 2 * n;
  off[21] = o; o += sizeof(double) * n;
  off[22] = o; o += sizeof(double) * nt;
  return (o + 15) / 16 * 16;
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

// ---- block reductions (NT threads) --------------------------------------------------------------------
template <int NT>
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
  for (int ww = 1; ww < (NT / 32); ++ww) {
    const double ov = red[2 * ww];
    const int oi = reinterpret_cast<int *>(red + 2 * ww + 1)[0];
    if (ov < val || (ov == val && oi >= 0 && (idx < 0 || oi < idx))) {
      val = ov;
      idx = oi;
    }
  }
}

template <int NT>
__device__ __forceinline__ double block_sum(double val, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = val;
  __syncthreads();
  double s = red[0];
#pragma unroll
  for (int ww = 1; ww < (NT / 32); ++ww) s += red[ww];
  return s;
}


__host__ __device__ inline size_t qp_scratch_span(int n, int nj, int OH, int qs = QP_QS) {  // bytes of the leading dead-during-gradient span
  const int tmax = nj * OH + n + 2;
  return sizeof(double) * ((size_t)qs * qs + 3 * n + 2 * (size_t)tmax + 3 * (size_t)(n + 2));
}

__device__ __forceinline__ QpView qp_view(unsigned char *smem_raw, int n, int nj, int OH, int m, int qs = QP_QS,
                                          int nt = QP_THREADS) {
  QpView s;
  size_t off[QP_NOFF];
  qp_smem_layout(n, nj, OH, m, off, qs, nt);
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
  return s;
}

struct QpDims {
  int n, nj, H, np, OH, m, has_vel, has_bnd;
  const double *G;     // Gram operator (L2 resident)
  const double *umax;  // MAX_input (shared memory copy)
  double *Mgl;         // this CTA's spill slab for the working-set inverse
  int ldg;
  double dt;
};

#define PF_START() do { if (prof && tid == 0) tck = clock64(); } while (0)
#define PF_ADD(k) do { if (prof && tid == 0) { const long long now_ = clock64(); pf[k] += now_ - tck; tck = now_; } } while (0)

// Solves   min 1/2 u'QQ u + ff'u  s.t. the rows described by (s.ocoef, s.orhs, lim, umax)   starting from the
// unconstrained minimiser whose primitives are in s.v0s / s.v.  On return (status 0) s.v holds the primitives of the
// optimum (controls in s.v[2n..3n)), s.lam / s.act / q the multipliers and the working set.
// status: 0 optimal, 2 infeasible, 3 numerical, 4 escalate (step_cap exceeded / working set outgrew QS).  Must be called by all NT threads of the CTA.
template <int NT, int QS>
__device__ __forceinline__ int qp_solve(const QpView &s, const QpDims &P, double cost0, double fupper, bool skip_solve,
                                        int &q_out, int &steps_out, int &qmax_seen, long long *pf, long long &tck,
                                        bool prof, int step_cap = 0x7fffffff, bool escalate_on_spill = false) {
  const int tid = threadIdx.x;
  const int n = P.n, nj = P.nj, H = P.H, np = P.np, OH = P.OH, m = P.m, has_vel = P.has_vel, has_bnd = P.has_bnd;
  const double *__restrict__ G = P.G;
  const double *umax = P.umax;
  double *Mgl = P.Mgl;
  const int ldg = P.ldg;
  const double dt = P.dt;
    int q = 0, status = -1, steps = 0;
    bool in_smem = true, polished = false;
    if (skip_solve) status = 0;
    double fval = cost0;
    const int max_steps = 20 * (m + n) + 100;
#define MAT(r_, c_) (in_smem ? s.Msm[(r_) + QS * (c_)] : Mgl[(r_) + (size_t)ldg * (c_)])
    while (status < 0) {
      // (0) primal recovery from the multipliers: v = v0 - G (C_W' lambda)
      if (q > 0) {
        const int T = s.toff[q];
        for (int t = tid; t < T; t += NT) s.twgt[t] = s.lam[s.towner[t]] * s.tcoef[t];
        __syncthreads();
        if (T <= QP_SMALL_T) {
          for (int base = 0; base < np; base += 6 * NT) {  // 6 primitives per thread, 2 terms per pass: 12 loads in flight
            double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            int t = 0;
            for (; t + 2 <= T; t += 2) {
              const double w0_ = s.twgt[t], w1_ = s.twgt[t + 1];
              const double *g0 = G + (size_t)s.trow[t] * np + base + tid, *g1 = G + (size_t)s.trow[t + 1] * np + base + tid;
              double l0[6], l1[6];
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                const bool ok = base + tid + j * NT < np;
                l0[j] = ok ? g0[j * NT] : 0.0;
                l1[j] = ok ? g1[j * NT] : 0.0;
              }
#pragma unroll
              for (int j = 0; j < 6; ++j) acc[j] += w0_ * l0[j] + w1_ * l1[j];
            }
            if (t < T) {
              const double w0_ = s.twgt[t];
              const double *g0 = G + (size_t)s.trow[t] * np + base + tid;
#pragma unroll
              for (int j = 0; j < 6; ++j)
                if (base + tid + j * NT < np) acc[j] += w0_ * g0[j * NT];
            }
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const int pi = base + tid + j * NT;
              if (pi < np) s.v[pi] = s.v0s[pi] - acc[j];
            }
          }
        } else {
          const double *__restrict__ Gu = G + 2 * n;  // control block of every primitive row
          for (int c = tid; c < n; c += NT) {
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
          for (int e = tid; e < n; e += NT) {  // B_theta u and B_omega u in closed form
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
      for (int cid = tid; cid < OH; cid += NT) {
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
      for (int e = tid; e < 2 * n; e += NT) {
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
      block_argmin<NT>(best, bidx, s.red);
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
          const int nch = NT / q > 0 ? NT / q : 1;  // chunks of columns per row, fixed summation order
          if (q <= NT) {
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
            for (int w = tid; w < q; w += NT) {
              const Desc dw = decode(s.act[w], OH, H, n, nj, s.ocoef);
              double acc = 0.0;
              for (int c = 0; c < q; ++c) acc += gram(dw, decode(s.act[c], OH, H, n, nj, s.ocoef), G, np) * s.lam[c];
              s.g[w] = viol_at_u0(s.act[w], OH, H, n, nj, s, umax) - acc;
            }
          }
          __syncthreads();
          for (int w = tid; w < q; w += NT) {
            double acc = 0.0;
            for (int c = 0; c < q; ++c) acc += MAT(w, c) * s.g[c];
            s.r[w] = acc;
          }
          __syncthreads();
          for (int w = tid; w < q; w += NT) s.lam[w] += s.r[w];
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
        if (steps > step_cap) {  // hand the problem to the heavy tier (k_fused.cu)
          status = 4;
          break;
        }
        // g_w = c_w QQ^-1 c_p'
        for (int w = tid; w < q; w += NT) s.g[w] = gram(decode(s.act[w], OH, H, n, nj, s.ocoef), dp, G, np);
        __syncthreads();
        // r = Minv g
        double part = 0.0;
        for (int w = tid; w < q; w += NT) {
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
        const double delta = sigma - block_sum<NT>(part, s.red);  // z'n+ in Goldfarb-Idnani's notation
        // t1: largest dual step keeping the multipliers non-negative
        double t1 = INFINITY;
        int l = -1;
        for (int w = tid; w < q; w += NT)
          if (s.r[w] > 0.0) {
            const double t = s.lam[w] / s.r[w];
            if (t < t1 || l < 0) {
              t1 = t;
              l = w;
            }
          }
        block_argmin<NT>(t1, l, s.red);
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
        for (int w = tid; w < q; w += NT) s.lam[w] -= t * s.r[w];
        lam_p += t;
        __syncthreads();
        PF_ADD(3);
        if (full) {
          if (in_smem && q + 1 > QS && escalate_on_spill) {
            status = 4;
            break;
          }
          if (in_smem && q + 1 > QS) {  // spill the inverse to the global slab
            for (int e = tid; e < q * q; e += NT) Mgl[(e % q) + (size_t)ldg * (e / q)] = s.Msm[(e % q) + QS * (e / q)];
            in_smem = false;
            __syncthreads();
          }
          // add p: bordered inverse  [[M + r r'/d, -r/d], [-r'/d, 1/d]]
          const double id = 1.0 / delta;
          {
            const int q1 = q + 1, tot = q1 * q1;
            for (int e0 = tid; e0 < tot; e0 += 4 * NT) {
              double old4[4];
              int rr4[4], cc4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = e0 + j * NT;
                rr4[j] = e % q1;
                cc4[j] = e / q1;
                old4[j] = (e < tot && rr4[j] < q && cc4[j] < q) ? MAT(rr4[j], cc4[j]) : 0.0;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = e0 + j * NT;
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
          for (int w = tid; w < q; w += NT) s.g[w] = MAT(w, l);
          __syncthreads();
          const double ip = 1.0 / s.g[l];
          for (int e0 = tid; e0 < q * q; e0 += 4 * NT) {
            double old4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = e0 + j * NT;
              old4[j] = e < q * q ? MAT(e % q, e / q) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = e0 + j * NT;
              if (e < q * q) MAT(e % q, e / q) = old4[j] - s.g[e % q] * s.g[e / q] * ip;
            }
          }
          __syncthreads();
          if (l != last) {
            for (int w = tid; w < q; w += NT) MAT(w, l) = MAT(w, last);
            __syncthreads();
            for (int w = tid; w < q; w += NT) MAT(l, w) = MAT(last, w);
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
          for (int w = tid; w < q; w += NT) {
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
    q_out = q;
  steps_out = steps;
  return status;
}

}  // namespace cfs
