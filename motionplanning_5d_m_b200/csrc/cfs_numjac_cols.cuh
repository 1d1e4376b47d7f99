// cfs_numjac_cols.cuh -- the num_jac evaluations of one trajectory as single-chain work items, for the fused persistent
// solver (k_fused.cu), where ONE problem's H waypoints must keep a whole CTA busy and the instruction footprint must stay
// small (the fused kernel is instruction-fetch bound with the one-thread-per-waypoint routine of cfs_numjac.cuh inlined:
// ncu stall "no_instruction" 12 of 20 cycles per issue).
//
// Same evaluated values as cfs_numjac.cuh (and hence as Lib/functions/num_jac.m applied to dist_arm_*).  num_jac never
// resets xp (num_jac.m:13-14), so column c is evaluated with joints < c at theta - eps/2, joint c at theta +- eps/2 and
// joints > c at theta.  Every evaluation of a waypoint is therefore a path in one small tree:
//     chain M ("all minus"):  links 0..NJ-1 at theta - eps/2; its running transforms Mm_l are the kinematic prefix of every
//                             column, its running minima pmin_l = min_{l' <= l} d_l' the distances of the unmoved links
//     chain B (base, y=f(x)): links 0..NJ-1 at theta
//     P_c  (yhi of column c): from Mm_{c-1}: link c at theta + eps/2, links > c at theta          (NJ - c link steps)
//     N_c  (ylo of column c): from Mm_c:     links > c at theta                                    (NJ - 1 - c link steps)
// => 2 NJ + NJ(NJ+1)/2 + NJ(NJ-1)/2 = 35 link steps per waypoint at NJ = 5 instead of 55, in two passes:
//     pass 1: H threads, chains (M_i, B_i) side by side; M stores Mm_1..Mm_{NJ-2} and pmin_0..pmin_{NJ-1} in shared memory
//     pass 2: NJ*H items of exactly NJ link steps each -- item (t, i) = P_t followed by N_{NJ-1-t} -- two per thread
// One loop body = one link transform + endpoints + distance for both chains (two independent FP64 dependency chains per
// thread): ~6 KB of SASS.
#pragma once
#include "cfs_geom.cuh"

namespace cfs {

// sin/cos of every waypoint's joint angles at theta (kind 0), theta+eps/2 (1), theta-eps/2 (2):
//   scw[((2*kind + 0)*NJ + k)*H + i] = cos,  scw[((2*kind + 1)*NJ + k)*H + i] = sin
template <int NJ, int NT>
__device__ __forceinline__ void numjac_sincos(const DevTables &tab, const double *xs, int H, double *scw) {
#pragma unroll 1
  for (int e = threadIdx.x; e < H * NJ; e += NT) {
    const int i = e / NJ, k = e - i * NJ;
    double s, c;
    sincos(xs[i * 2 * NJ + k] + tab.link[k].th_off, &s, &c);
    const double cc = c * CFS_NUMJAC_COSH, ss = s * CFS_NUMJAC_COSH;  // angle addition, see cfs_types.cuh
    scw[(0 * NJ + k) * H + i] = c;
    scw[(1 * NJ + k) * H + i] = s;
    scw[(2 * NJ + k) * H + i] = fma(-s, CFS_NUMJAC_SINH, cc);  // cos(theta + eps/2)   (num_jac.m:11)
    scw[(3 * NJ + k) * H + i] = fma(c, CFS_NUMJAC_SINH, ss);   // sin(theta + eps/2)
    scw[(4 * NJ + k) * H + i] = fma(s, CFS_NUMJAC_SINH, cc);   // cos(theta - eps/2)   (num_jac.m:13)
    scw[(5 * NJ + k) * H + i] = fma(-c, CFS_NUMJAC_SINH, ss);  // sin(theta - eps/2)
  }
}

// shared-memory scratch of the gradient phase, in doubles (aliases the QP scratch span, see k_fused.cu)
__host__ __device__ inline size_t numjac_scratch_doubles(int nj, int H, int OH) {
  return (size_t)6 * nj * H            // scw
         + (size_t)2 * OH * nj         // fv: f(x+), f(x-) per (obstacle, waypoint, column)
         + (size_t)12 * (nj > 2 ? nj - 2 : 0) * H  // pm: Mm_1 .. Mm_{NJ-2}
         + (size_t)OH * nj;            // pmin
}

struct NumjacScratch {
  double *scw, *fv, *pm, *pmin;
};
__device__ __forceinline__ NumjacScratch numjac_scratch(double *base, int nj, int H, int OH) {
  NumjacScratch w;
  w.scw = base;
  w.fv = w.scw + 6 * nj * H;
  w.pm = w.fv + 2 * OH * nj;
  w.pmin = w.pm + 12 * (nj > 2 ? nj - 2 : 0) * H;
  return w;
}

// One work item = NJ link steps.  pass 1: type 0 = chain B, type 1 = chain M.  pass 2: type t in [0, NJ): P_t then
// N_{NJ-1-t}.  Step s of an item: link l, sin/cos kind, whether the step starts a segment, whether it ends one.
struct StepDesc {
  int l, kind;
  int start;  // 0 continue, 1 segment starts here
  int end;    // 0 no; 1 write base distance (orhs); 2 write fv[...][col][sign]; pass-1 chain M handles its own stores
  int col, sign;
};
template <int NJ>
__device__ __forceinline__ StepDesc step_desc(int pass, int type, int s) {
  StepDesc d;
  if (pass == 1) {
    d.l = s;
    d.kind = type ? 2 : 0;
    d.start = (s == 0);
    d.end = (s == NJ - 1) ? (type ? 2 : 1) : 0;
    d.col = NJ - 1;  // chain M is also ylo of the last column
    d.sign = 1;
  } else {
    const int np = NJ - type;  // link steps of P_type
    if (s < np) {
      d.l = type + s;
      d.kind = (s == 0) ? 1 : 0;
      d.start = (s == 0);
      d.end = (s == np - 1) ? 2 : 0;
      d.col = type;
      d.sign = 0;
    } else {
      d.l = s;
      d.kind = 0;
      d.start = (s == np);
      d.end = (s == NJ - 1) ? 2 : 0;
      d.col = NJ - 1 - type;
      d.sign = 1;
    }
  }
  return d;
}

// Two items side by side against obstacles j0, j0+1.  Item X: waypoint iX, type tyX.  Chain B is skipped when !hasB.
// Results go to w.fv / orhs_base (the base distance, margin not yet subtracted) / w.pm / w.pmin.
template <int NJ>
__device__ __forceinline__ void numjac_items2(const DevTables &tab, const NumjacScratch &w, int H, int pass, int iA, int tyA,
                                              int iB, int tyB, bool hasB, int j0, int nobs, int &touched, double *dbase) {
  const double *scw = w.scw;
  Xf MA, MB;
  double pA[6], pB[6];
  double dA[2], dB[2];
  dA[0] = dA[1] = dB[0] = dB[1] = INFINITY;
#pragma unroll 1
  for (int s = 0; s < NJ; ++s) {
    const StepDesc a = step_desc<NJ>(pass, tyA, s), b = step_desc<NJ>(pass, tyB, s);
    const double cA = scw[((2 * a.kind + 0) * NJ + a.l) * H + iA], sA = scw[((2 * a.kind + 1) * NJ + a.l) * H + iA];
    const double cB = scw[((2 * b.kind + 0) * NJ + b.l) * H + iB], sB = scw[((2 * b.kind + 1) * NJ + b.l) * H + iB];
    // ---- segment start: kinematic prefix Mm_{l-1} and the minima of the links it already passed ----
    if (a.start && a.l > 0) {
      if (a.l == 1) {
        xf_first(tab.link[0], scw[((2 * 2 + 0) * NJ + 0) * H + iA], scw[((2 * 2 + 1) * NJ + 0) * H + iA], MA);
      } else {
        const double *src = w.pm + ((size_t)iA * (NJ - 2) + (a.l - 2)) * 12;
#pragma unroll
        for (int e = 0; e < 12; ++e) MA.m[e] = src[e];
      }
      dA[0] = w.pmin[((size_t)j0 * H + iA) * NJ + a.l - 1];
      if (j0 + 1 < nobs) dA[1] = w.pmin[((size_t)(j0 + 1) * H + iA) * NJ + a.l - 1];
    }
    if (b.start && b.l > 0) {
      if (b.l == 1) {
        xf_first(tab.link[0], scw[((2 * 2 + 0) * NJ + 0) * H + iB], scw[((2 * 2 + 1) * NJ + 0) * H + iB], MB);
      } else {
        const double *src = w.pm + ((size_t)iB * (NJ - 2) + (b.l - 2)) * 12;
#pragma unroll
        for (int e = 0; e < 12; ++e) MB.m[e] = src[e];
      }
      dB[0] = w.pmin[((size_t)j0 * H + iB) * NJ + b.l - 1];
      if (j0 + 1 < nobs) dB[1] = w.pmin[((size_t)(j0 + 1) * H + iB) * NJ + b.l - 1];
    }
    // ---- one link step of both chains ----
    if (a.l == 0)
      xf_first(tab.link[0], cA, sA, MA);
    else
      xf_step_inplace(MA, tab.link[a.l], cA, sA);
    if (b.l == 0)
      xf_first(tab.link[0], cB, sB, MB);
    else
      xf_step_inplace(MB, tab.link[b.l], cB, sB);
    link_endpoints(MA, tab.link[a.l], tab.base, pA);
    link_endpoints(MB, tab.link[b.l], tab.base, pB);
#pragma unroll 1
    for (int jj = 0; jj < 2; ++jj)
      if (j0 + jj < nobs) {
        const double da = link_obs_key(pA, tab.obs[j0 + jj], touched);  // keys (signed squares), see cfs_geom.cuh
        int tb = 0;
        const double db = link_obs_key(pB, tab.obs[j0 + jj], tb);
        if (hasB) touched |= tb;
        // strict <: the first minimal link wins (dist_arm_3D_Heu_2.m:25-28)
        if (jj == 0) {
          dA[0] = da < dA[0] ? da : dA[0];
          dB[0] = db < dB[0] ? db : dB[0];
        } else {
          dA[1] = da < dA[1] ? da : dA[1];
          dB[1] = db < dB[1] ? db : dB[1];
        }
      }
    // ---- stores ----
#pragma unroll 1
    for (int ch = 0; ch < (hasB ? 2 : 1); ++ch) {
      const StepDesc &d = ch ? b : a;
      const int i = ch ? iB : iA, ty = ch ? tyB : tyA;
      const double d0 = ch ? dB[0] : dA[0], d1 = ch ? dB[1] : dA[1];
      if (pass == 1 && ty == 1) {  // chain M: running minima and kinematic prefixes for pass 2
        w.pmin[((size_t)j0 * H + i) * NJ + d.l] = d0;
        if (j0 + 1 < nobs) w.pmin[((size_t)(j0 + 1) * H + i) * NJ + d.l] = d1;
        if (NJ > 2 && d.l >= 1 && d.l <= NJ - 2 && j0 == 0) {
          double *dst = w.pm + ((size_t)i * (NJ - 2) + (d.l - 1)) * 12;
          const Xf &M = ch ? MB : MA;
#pragma unroll
          for (int e = 0; e < 12; ++e) dst[e] = M.m[e];
        }
      }
      if (d.end) {  // one evaluation of dist_arm is complete: the only square roots of the chain
        const double e0 = key_to_dist(d0), e1 = key_to_dist(d1);
        if (d.end == 1) {
          dbase[j0 * H + i] = e0;
          if (j0 + 1 < nobs) dbase[(j0 + 1) * H + i] = e1;
        } else {
          w.fv[(((size_t)j0 * H + i) * NJ + d.col) * 2 + d.sign] = e0;
          if (j0 + 1 < nobs) w.fv[(((size_t)(j0 + 1) * H + i) * NJ + d.col) * 2 + d.sign] = e1;
        }
      }
    }
  }
}

}  // namespace cfs
