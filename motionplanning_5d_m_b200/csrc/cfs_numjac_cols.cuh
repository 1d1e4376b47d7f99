// cfs_numjac_cols.cuh -- the num_jac evaluations of one trajectory as H*(2NJ+1) single-chain work items, for the fused
// persistent solver (k_fused.cu), where ONE problem's H waypoints must keep a whole CTA busy and the instruction
// footprint must stay small (the fused kernel is instruction-fetch bound: ncu stall "no_instruction" 12 of 20 cycles
// per issue with the one-thread-per-waypoint routine inlined).
//
// Same evaluated values as cfs_numjac.cuh (and hence as Lib/functions/num_jac.m applied to dist_arm_*): item
// (i, col, sign) evaluates f at x with joints < col at theta - eps/2 (num_jac never resets xp, num_jac.m:13-14), joint
// col at theta +- eps/2 and joints > col at theta; the extra item per waypoint is the base evaluation y = f(x).
// One loop body = one link transform + endpoints + distance: ~6 KB of SASS instead of ~90 KB.
#pragma once
#include "cfs_geom.cuh"

namespace cfs {

// sin/cos of every waypoint's joint angles at theta (kind 0), theta+eps/2 (1), theta-eps/2 (2):
//   scw[((2*kind + 0)*NJ + k)*H + i] = cos,  scw[((2*kind + 1)*NJ + k)*H + i] = sin
template <int NJ, int NT>
__device__ __forceinline__ void numjac_sincos(const DevTables &tab, const double *xs, int H, double *scw) {
  const double hh = CFS_NUMJAC_EPS / 2;  // num_jac.m:11,13
#pragma unroll 1
  for (int e = threadIdx.x; e < 3 * H * NJ; e += NT) {
    const int kind = e / (H * NJ), r = e - kind * (H * NJ);
    const int i = r / NJ, k = r - i * NJ;
    const double th = xs[i * 2 * NJ + k];
    const double arg = (kind == 0 ? th : (kind == 1 ? th + hh : th - hh)) + tab.link[k].th_off;
    double s, c;
    sincos(arg, &s, &c);
    scw[((2 * kind + 0) * NJ + k) * H + i] = c;
    scw[((2 * kind + 1) * NJ + k) * H + i] = s;
  }
}

// Two evaluation chains side by side (two independent FP64 dependency chains per thread: the chain is latency bound),
// each against obstacles j0, j0+1.  Chain X: waypoint iX, column colX (colX == NJ: base evaluation, all joints at theta),
// signX 0: joint colX at +eps/2, 1: at -eps/2.  Chain B is skipped when !hasB.
template <int NJ>
__device__ __forceinline__ void numjac_chain2(const DevTables &tab, const double *scw, int H, int iA, int colA, int signA,
                                              int iB, int colB, int signB, bool hasB, int j0, int nobs, int &touched,
                                              double dA[2], double dB[2]) {
  dA[0] = dA[1] = dB[0] = dB[1] = INFINITY;
  Xf MA, MB;
  double pA[6], pB[6];
#pragma unroll 1
  for (int l = 0; l < NJ; ++l) {
    // joints < col: theta - eps/2 ; joint col: +-eps/2 ; joints > col (and the base evaluation): theta
    const int kA = (colA == NJ || l > colA) ? 0 : ((l < colA || signA) ? 2 : 1);
    const int kB = (colB == NJ || l > colB) ? 0 : ((l < colB || signB) ? 2 : 1);
    const double cA = scw[((2 * kA + 0) * NJ + l) * H + iA], sA = scw[((2 * kA + 1) * NJ + l) * H + iA];
    const double cB = scw[((2 * kB + 0) * NJ + l) * H + iB], sB = scw[((2 * kB + 1) * NJ + l) * H + iB];
    if (l == 0) {
      xf_first(tab.link[0], cA, sA, MA);
      xf_first(tab.link[0], cB, sB, MB);
    } else {
      xf_step_inplace(MA, tab.link[l], cA, sA);
      xf_step_inplace(MB, tab.link[l], cB, sB);
    }
    link_endpoints(MA, tab.link[l], tab.base, pA);
    link_endpoints(MB, tab.link[l], tab.base, pB);
#pragma unroll 1
    for (int jj = 0; jj < 2; ++jj)
      if (j0 + jj < nobs) {
        const double a = link_obs_dist(pA, tab.obs[j0 + jj], touched);
        int tb = 0;
        const double b = link_obs_dist(pB, tab.obs[j0 + jj], tb);
        if (hasB) touched |= tb;
        // strict <: the first minimal link wins (dist_arm_3D_Heu_2.m:25-28)
        if (jj == 0) {
          dA[0] = a < dA[0] ? a : dA[0];
          dB[0] = b < dB[0] ? b : dB[0];
        } else {
          dA[1] = a < dA[1] ? a : dA[1];
          dB[1] = b < dB[1] ? b : dB[1];
        }
      }
  }
}

}  // namespace cfs
