// cfs_numjac.cuh -- one waypoint of CFS_FANUC.get_con's inner loop body (Lib/CFS_FANUC.m:113-118):
//     [distance,linkid] = dist_arm_all(theta,...);  Diff = num_jac(f,theta)
// as a device routine shared by the stand-alone K1 kernel (k_grad.cu) and the fused persistent solver (k_fused.cu).
// See k_grad.cu for how the 11 evaluations of num_jac are restructured without changing any evaluated value.
#pragma once
#include "cfs_geom.cuh"

namespace cfs {

#define GRAD_THREADS 128

__device__ __forceinline__ double min_first(double cur, double cand) { return cand < cur ? cand : cur; }

// sc: this CTA's sin/cos cache [6 kinds][NJ][NT threads] in shared memory (kinds: c0,s0 at theta; cp,sp at
// theta+eps/2; cm,sm at theta-eps/2), column `tid` belongs to the calling thread.  thp: the NJ joint angles.
// out.grad(j, k, value) receives Diff(k) for obstacle j, out.dist(j, distance, linkid) the base evaluation.
template <int NJ, int OC, int NT, class Out>
__device__ __forceinline__ void numjac_waypoint(const DevTables &tab, double (*sc)[NJ][NT], double (*pmS)[NT], int tid,
                                                const double *thp, int nobs, int &touched, Out &out) {
#pragma unroll
  for (int k = 0; k < NJ; ++k) {
    const double th = thp[k];
    const double off = tab.link[k].th_off;
    double s, c;
    sincos(th + off, &s, &c);
    const double cc = c * CFS_NUMJAC_COSH, ss = s * CFS_NUMJAC_COSH;
    sc[0][k][tid] = c;
    sc[1][k][tid] = s;
    sc[2][k][tid] = fma(-s, CFS_NUMJAC_SINH, cc);  // cos(theta + eps/2)   (num_jac.m:11)
    sc[3][k][tid] = fma(c, CFS_NUMJAC_SINH, ss);   // sin(theta + eps/2)
    sc[4][k][tid] = fma(s, CFS_NUMJAC_SINH, cc);   // cos(theta - eps/2)   (num_jac.m:13)
    sc[5][k][tid] = fma(-c, CFS_NUMJAC_SINH, ss);  // sin(theta - eps/2)
  }

  for (int j0 = 0; j0 < nobs; j0 += OC) {
    double dbase[OC], dpre[OC];
    int lid[OC];
#pragma unroll
    for (int jj = 0; jj < OC; ++jj) {
      dbase[jj] = INFINITY;
      dpre[jj] = INFINITY;
      lid[jj] = 0;
    }
    Xf M;  // the running prefix M_1^- ... M_k^- lives in shared memory (pmS[12][thread]): 24 registers less
    double p[6];
    // ---- y = f(x): base evaluation, gives distance and linkid (CFS_FANUC.m:115) ----
#pragma unroll 1
    for (int l = 0; l < NJ; ++l) {
      if (l == 0) {
        xf_first(tab.link[0], sc[0][0][tid], sc[1][0][tid], M);
      } else {
        xf_step_inplace(M, tab.link[l], sc[0][l][tid], sc[1][l][tid]);
      }
      link_endpoints(M, tab.link[l], tab.base, p);
#pragma unroll
      for (int jj = 0; jj < OC; ++jj)
        if (j0 + jj < nobs) {
          const double d = link_obs_key(p, tab.obs[j0 + jj], touched);  // keys (signed squares), see cfs_geom.cuh
          if (d < dbase[jj]) {  // strict <: first minimal link (dist_arm_3D_Heu_2.m:25-28)
            dbase[jj] = d;
            lid[jj] = l + 1;
          }
        }
    }
    // ---- columns of num_jac ----
#pragma unroll 1
    for (int k = 0; k < NJ; ++k) {
      double dpl[OC], dmi[OC], dk[OC];
      // yhi = f(xp), xp(k) = x(k)+eps/2, joints < k at x-eps/2
      if (k == 0) {
        xf_first(tab.link[0], sc[2][0][tid], sc[3][0][tid], M);
      } else {
#pragma unroll
        for (int e = 0; e < 12; ++e) M.m[e] = pmS[e][tid];
        xf_step_inplace(M, tab.link[k], sc[2][k][tid], sc[3][k][tid]);
      }
      link_endpoints(M, tab.link[k], tab.base, p);
#pragma unroll
      for (int jj = 0; jj < OC; ++jj)
        dpl[jj] = (j0 + jj < nobs) ? min_first(dpre[jj], link_obs_key(p, tab.obs[j0 + jj], touched)) : 0.0;
#pragma unroll 1
      for (int l = k + 1; l < NJ; ++l) {
        xf_step_inplace(M, tab.link[l], sc[0][l][tid], sc[1][l][tid]);
        link_endpoints(M, tab.link[l], tab.base, p);
#pragma unroll
        for (int jj = 0; jj < OC; ++jj)
          if (j0 + jj < nobs) dpl[jj] = min_first(dpl[jj], link_obs_key(p, tab.obs[j0 + jj], touched));
      }
      // ylo = f(xp), xp(k) = x(k)-eps/2
      if (k == 0) {
        xf_first(tab.link[0], sc[4][0][tid], sc[5][0][tid], M);
      } else {
#pragma unroll
        for (int e = 0; e < 12; ++e) M.m[e] = pmS[e][tid];
        xf_step_inplace(M, tab.link[k], sc[4][k][tid], sc[5][k][tid]);
      }
      if (k + 1 < NJ) {  // running prefix M_1^- ... M_k^-
#pragma unroll
        for (int e = 0; e < 12; ++e) pmS[e][tid] = M.m[e];
      }
      link_endpoints(M, tab.link[k], tab.base, p);
#pragma unroll
      for (int jj = 0; jj < OC; ++jj) {
        dk[jj] = (j0 + jj < nobs) ? link_obs_key(p, tab.obs[j0 + jj], touched) : 0.0;
        dmi[jj] = min_first(dpre[jj], dk[jj]);
      }
#pragma unroll 1
      for (int l = k + 1; l < NJ; ++l) {
        xf_step_inplace(M, tab.link[l], sc[0][l][tid], sc[1][l][tid]);
        link_endpoints(M, tab.link[l], tab.base, p);
#pragma unroll
        for (int jj = 0; jj < OC; ++jj)
          if (j0 + jj < nobs) dmi[jj] = min_first(dmi[jj], link_obs_key(p, tab.obs[j0 + jj], touched));
      }
#pragma unroll
      for (int jj = 0; jj < OC; ++jj)
        if (j0 + jj < nobs) {
          out.grad(j0 + jj, k, (key_to_dist(dpl[jj]) - key_to_dist(dmi[jj])) / CFS_NUMJAC_EPS);  // num_jac.m:15
          dpre[jj] = min_first(dpre[jj], dk[jj]);
        }
    }
#pragma unroll
    for (int jj = 0; jj < OC; ++jj)
      if (j0 + jj < nobs) {
        out.dist(j0 + jj, key_to_dist(dbase[jj]), lid[jj]);
      }
  }
}

}  // namespace cfs
