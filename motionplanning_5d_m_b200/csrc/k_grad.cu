// k_grad.cu -- K1 (num_jac), K1d (DERIVEST) distance/gradient kernels and K6 (RRT node checks), sm_100a.
//
// K1  replaces CFS_FANUC.get_con's inner loop body, Lib/CFS_FANUC.m:113-118:
//        [distance,linkid] = dist_arm_all(theta,...);  Diff = num_jac(f,theta)
//     One thread per (problem, waypoint).  The 11 evaluations of num_jac (Lib/functions/num_jac.m:2,10-16)
//     are restructured, without changing any evaluated value, around two facts:
//       * num_jac never resets xp(i) after column i (num_jac.m:13-14), so column k is evaluated with joints
//         1..k-1 at theta-eps/2.  The kinematic prefix M_1^- ... M_{k-1}^- is therefore a running product
//         shared by every later column, and the distances of links < k are the ones the "minus" evaluation
//         of their own column already produced.
//       * only links >= k move when joint k is perturbed.
//     => 15 sincos + 35 link transforms/distances per waypoint instead of 55 + 55, all in registers;
//     the per-thread sin/cos cache lives in shared memory (conflict-free [value][thread] layout).
// K1d replaces the script path M16iB/main_CFS.m:231-237: derivest(@(x) dist_link_Heu(...,linkid), theta(s)).
// K6  replaces RRT_FANUC.feasible (Lib/RRT_FANUC.m:146-181) and the nearest/steer scan (:116-129).
#include "cfs_geom.cuh"
#include "cfs_kernels.cuh"
#include "cfs_numjac.cuh"

namespace cfs {

#ifndef GRAD_MINB
#define GRAD_MINB 4  // 4 x 128 threads per SM: <= 128 registers per thread
#endif

// ============================================================================================================
// K1: num_jac
// ============================================================================================================
template <int NJ, int OC>
__global__ void __launch_bounds__(GRAD_THREADS, GRAD_MINB) k_grad_numjac(GradArgs a) {
  __shared__ alignas(128) DevTables tab;
  __shared__ alignas(8) uint64_t mbar;
  // dynamic shared memory: sin/cos cache [6 kinds][NJ][thread] (kinds: c0,s0 (theta), cp,sp (theta+eps/2), cm,sm
  // (theta-eps/2)) followed by the running "all minus" kinematic prefix [12][thread] of every thread's waypoint
  extern __shared__ __align__(16) double k1_dyn[];
  double (*sc)[NJ][GRAD_THREADS] = reinterpret_cast<double (*)[NJ][GRAD_THREADS]>(k1_dyn);
  double (*pmS)[GRAD_THREADS] = reinterpret_cast<double (*)[GRAD_THREADS]>(k1_dyn + 6 * NJ * GRAD_THREADS);

  const int count = a.count ? *a.count : a.nslots;
  const long long total = (long long)count * a.H;
  if ((long long)blockIdx.x * blockDim.x >= total) return;  // whole CTA idle
  tma_stage(&tab, a.tab, tab_bytes(a.nobs), &mbar);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int tid = threadIdx.x;
  const int slot = (int)(t / a.H), i = (int)(t % a.H);
  const int prob = a.list ? a.list[slot] : slot;
  const double *thp = a.x + prob * a.ld_prob + i * a.ld_i;

  int touched = 0;
  struct Sink {
    const GradArgs &a;
    long long prob;
    int i;
    __device__ __forceinline__ long long idx(int j) const { return prob * a.o_prob + (long long)j * a.o_obs + i * a.o_i; }
    __device__ __forceinline__ void grad(int j, int k, double v) const { a.grad[idx(j) * NJ + k] = v; }
    __device__ __forceinline__ void dist(int j, double d, int lid) const {
      a.dist[idx(j)] = d;
      if (a.linkid) a.linkid[idx(j)] = lid;
    }
  } out{a, prob, i};
  numjac_waypoint<NJ, OC, GRAD_THREADS>(tab, sc, pmS, tid, thp, a.nobs, touched, out);
  if (touched && a.flags) atomicOr(&a.flags[prob], 0x100);
}

template <int NJ>
static cudaError_t launch_numjac_nj(const GradArgs &a, cudaStream_t s) {
  const long long total = (long long)a.nslots * a.H;
  if (total <= 0) return cudaSuccess;
  const int grid = (int)((total + GRAD_THREADS - 1) / GRAD_THREADS);
  const size_t dyn = sizeof(double) * (6 * NJ + 12) * GRAD_THREADS;
  if (a.nobs <= 1) {
    if (dyn > 48 * 1024) cudaFuncSetAttribute(k_grad_numjac<NJ, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    k_grad_numjac<NJ, 1><<<grid, GRAD_THREADS, dyn, s>>>(a);
  } else {
    if (dyn > 48 * 1024) cudaFuncSetAttribute(k_grad_numjac<NJ, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    k_grad_numjac<NJ, 2><<<grid, GRAD_THREADS, dyn, s>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_grad_numjac(const GradArgs &a, cudaStream_t s) {
  switch (a.nj) {
    case 2: return launch_numjac_nj<2>(a, s);
    case 3: return launch_numjac_nj<3>(a, s);
    case 4: return launch_numjac_nj<4>(a, s);
    case 5: return launch_numjac_nj<5>(a, s);
    case 6: return launch_numjac_nj<6>(a, s);
    default: return cudaErrorInvalidValue;
  }
}

// ============================================================================================================
// K1d: DERIVEST gradients of dist_link(linkid)
//   one thread per (problem, waypoint, obstacle, joint s)
// ============================================================================================================
#ifndef K1D_MINB
#define K1D_MINB 3
#endif
// "e precedes c" in MATLAB's ascending sort of der_romb (derivest.m:442): smaller value first, NaNs last, stable
__device__ __forceinline__ bool dv_precedes(double ve, int e, double vc, int c) {
  const bool lt = ve < vc || (isnan(vc) && !isnan(ve));
  const bool eq = (ve == vc) || (isnan(vc) && isnan(ve));
  return lt || (eq && e < c);
}

// K1d, two phases inside one CTA of 128 (waypoint, obstacle) pairs:
//   phase 1  one thread per pair: sin/cos of the joints, base evaluation (distance, linkid: M16iB/main_CFS.m:229), and the
//            CTA-wide list of work items (pair, joint s) with s < linkid -- dist_link(linkid) does not depend on joints
//            >= linkid, so every f_del of those joints is exactly 0 and derivest returns 0 (45 % of the reference line's
//            waypoints have linkid 1: a thread-per-joint mapping leaves half of the lanes idle);
//   phase 2  one thread per work item, dense lanes: the 26 x 2 evaluations f(x0 +- h*delta_k) (derivest.m:366-371) with both
//            signs side by side (two independent dependency chains), then der_init / rombextrap / trimmed selection.
// Register diet (234 -> 168 registers, 2 -> 3 CTAs/SM): the 23 derivative estimates, the 19 Romberg extrapolants and their
// error estimates live in shared-memory columns of the thread (der_init overwrites f_del in place, der_romb overwrites
// der_init in place), and the trimmed-sort selection finds the two smallest / two largest extrapolants by four argmin passes
// instead of ranking all 19.  Every evaluated value and every operation order is unchanged.
template <int NJ>
__global__ void __launch_bounds__(GRAD_THREADS, K1D_MINB) k_grad_derivest(GradArgs a, int pairs_per_cta) {
  __shared__ alignas(128) DevTables tab;
  __shared__ alignas(128) DerivestTab dv;
  __shared__ alignas(8) uint64_t mbar;
  __shared__ int lid_s[GRAD_THREADS], off_s[GRAD_THREADS + 1], wsum[GRAD_THREADS / 32];
  __shared__ long long o_s[GRAD_THREADS];
  __shared__ unsigned short items[GRAD_THREADS * NJ];
  extern __shared__ __align__(16) double k1d_dyn[];
  double (*fdel_s)[GRAD_THREADS] = reinterpret_cast<double (*)[GRAD_THREADS]>(k1d_dyn);  // f_del -> der_init -> der_romb
  double (*err_s)[GRAD_THREADS] = fdel_s + DV_NDEL;                                      // errest of every extrapolant
  double (*cs_s)[GRAD_THREADS] = err_s + DV_NEST;                                        // cos (rows 0..NJ-1), sin (NJ..2NJ-1)
  double (*th_s)[GRAD_THREADS] = cs_s + 2 * NJ;                                          // joint angles of the pair

  const int count = a.count ? *a.count : a.nslots;
  const long long total = (long long)count * a.H * a.nobs;  // pairs
  if ((long long)blockIdx.x * pairs_per_cta >= total) return;
  tma_stage(&tab, a.tab, tab_bytes(a.nobs), &mbar);
  for (int w = threadIdx.x; w < (int)(sizeof(DerivestTab) / sizeof(double)); w += blockDim.x)
    reinterpret_cast<double *>(&dv)[w] = reinterpret_cast<const double *>(a.dv)[w];
  __syncthreads();
  const int tid = threadIdx.x;
  // pairs_per_cta <= 128 pairs per CTA: small launches (late outer iterations) use fewer pairs per CTA so that phase 2 is one
  // round of items instead of up to NJ
  const long long t = (long long)blockIdx.x * pairs_per_cta + tid;
  int touched = 0, lid = 0, prob = -1;
  Xf M;
  double p[6];
  // ---- phase 1 ----
  if (tid < pairs_per_cta && t < total) {
    const int j = (int)(t % a.nobs);
    long long rest = t / a.nobs;
    const int i = (int)(rest % a.H);
    const int slot = (int)(rest / a.H);
    prob = a.list ? a.list[slot] : slot;
    const double *thp = a.x + prob * a.ld_prob + i * a.ld_i;
    double dbase = INFINITY;
#pragma unroll 1
    for (int l = 0; l < NJ; ++l) {
      const double thl = thp[l];
      double sn, cs;
      sincos(thl + tab.link[l].th_off, &sn, &cs);
      th_s[l][tid] = thl;
      cs_s[l][tid] = cs;
      cs_s[NJ + l][tid] = sn;
      if (l == 0)
        xf_first(tab.link[0], cs, sn, M);
      else
        xf_step_inplace(M, tab.link[l], cs, sn);
      if (a.no_off) continue;
      link_endpoints(M, tab.link[l], tab.base, p);
      const double d = link_obs_dist(p, tab.obs[j], touched);
      if (d < dbase) {
        dbase = d;
        lid = l + 1;
      }
    }
    const long long o = prob * a.o_prob + (long long)j * a.o_obs + i * a.o_i;
    if (a.no_off) {  // CHOMP_FANUC.dm_f: the same chain without the joint offsets; every link's distance is kept
#pragma unroll 1
      for (int l = 0; l < NJ; ++l) {
        double sn, cs;
        sincos(thp[l], &sn, &cs);
        if (l == 0)
          xf_first(tab.link[0], cs, sn, M);
        else
          xf_step_inplace(M, tab.link[l], cs, sn);
        link_endpoints(M, tab.link[l], tab.base, p);
        const double d = link_obs_dist(p, tab.obs[j], touched);
        if (a.linkdist) a.linkdist[o * NJ + l] = d;
        if (d < dbase) {
          dbase = d;
          lid = l + 1;
        }
      }
    }
    o_s[tid] = o;
    a.dist[o] = dbase;
    if (a.linkid) a.linkid[o] = lid;
    for (int k = lid; k < NJ; ++k) a.grad[o * NJ + k] = 0.0;
  }
  lid_s[tid] = lid;
  // exclusive scan of the item counts over the CTA
  {
    int inc = lid;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, d);
      if ((tid & 31) >= d) inc += v;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
    off_s[tid] = base + inc - lid;
    if (tid == GRAD_THREADS - 1) off_s[GRAD_THREADS] = base + inc;
  }
  __syncthreads();
  for (int k = 0; k < lid; ++k) items[off_s[tid] + k] = (unsigned short)(tid * 8 + k);
  __syncthreads();
  const int nitems = off_s[GRAD_THREADS];
  // ---- phase 2 ----
#pragma unroll 1
  for (int it = tid; it < nitems; it += GRAD_THREADS) {
    const int pr = items[it] >> 3, s = items[it] & 7, lidp = lid_s[pr];
    const long long tp = (long long)blockIdx.x * pairs_per_cta + pr;
    const ObsTab &ob = tab.obs[(int)(tp % a.nobs)];
    const double ths = th_s[s][pr], offs = tab.link[s].th_off;
    Xf Ps, MB;  // Ps: product of links < s (unused when s == 0)
#pragma unroll 1
    for (int l = 0; l < s; ++l) {
      if (l == 0)
        xf_first(tab.link[0], cs_s[0][pr], cs_s[NJ][pr], Ps);
      else
        xf_step_inplace(Ps, tab.link[l], cs_s[l][pr], cs_s[NJ + l][pr]);
    }
    const double h = ths > 0.02 ? ths : 0.02;  // par.NominalStep = max(x0,0.02)  derivest.m:229
#pragma unroll 1
    for (int kk = 0; kk < DV_NDEL; ++kk) {
      const double step = h * dv.delta[kk];
      double snA, csA, snB, csB;
      sincos((ths + step) + offs, &snA, &csA);  // derivest.m:369
      sincos((ths - step) + offs, &snB, &csB);  // derivest.m:370
      if (s == 0) {
        xf_first(tab.link[0], csA, snA, M);
        xf_first(tab.link[0], csB, snB, MB);
      } else {
        xf_step(Ps, tab.link[s], csA, snA, M);
        xf_step(Ps, tab.link[s], csB, snB, MB);
      }
#pragma unroll 1
      for (int l = s + 1; l < lidp; ++l) {
        const double cl = cs_s[l][pr], sl = cs_s[NJ + l][pr];
        xf_step_inplace(M, tab.link[l], cl, sl);
        xf_step_inplace(MB, tab.link[l], cl, sl);
      }
      double pB[6];
      link_endpoints(M, tab.link[lidp - 1], tab.base, p);
      link_endpoints(MB, tab.link[lidp - 1], tab.base, pB);
      const double fA = link_obs_dist(p, ob, touched);
      const double fB = link_obs_dist(pB, ob, touched);
      fdel_s[kk][tid] = (fA - fB) / 2;  // odd transformation, derivest.m:376
    }
    // der_init (derivest.m:415-418), in place: entry e only reads entries e and e+1
#pragma unroll 1
    for (int e = 0; e < DV_NE; ++e)
      fdel_s[e][tid] = (fdel_s[e][tid] * dv.fdarule[0] + fdel_s[e + 1][tid] * dv.fdarule[1]) / (h * dv.delta[e]);
    // rombextrap (derivest.m:512-526), in place: extrapolant c reads der_init entries c..c+3 and replaces entry c
#pragma unroll 1
    for (int c = 0; c < DV_NEST; ++c) {
      double di[4], qtr[3], coef[3];
#pragma unroll
      for (int r = 0; r < 4; ++r) di[r] = fdel_s[c + r][tid];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) acc += dv.q[r][cc] * di[r];
        qtr[cc] = acc;
      }
      coef[2] = qtr[2] / dv.rr[2][2];
      coef[1] = (qtr[1] - dv.rr[1][2] * coef[2]) / dv.rr[1][1];
      coef[0] = ((qtr[0] - dv.rr[0][1] * coef[1]) - dv.rr[0][2] * coef[2]) / dv.rr[0][0];
      double ss = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double res = di[r] - ((dv.rmat[r][0] * coef[0] + dv.rmat[r][1] * coef[1]) + dv.rmat[r][2] * coef[2]);
        ss += res * res;
      }
      fdel_s[c][tid] = coef[0];
      err_s[c][tid] = sqrt(ss) * dv.errfac;
    }
    // sort ascending (stable), drop ranks {1,2,nest-1,nest}, take the minimum error (derivest.m:442-462): the dropped
    // extrapolants are the two that precede all others and the two that all others precede
    int drop[4] = {-1, -1, -1, -1};
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
      int bi = -1;
      double bv = 0.0;
#pragma unroll 1
      for (int c = 0; c < DV_NEST; ++c) {
        if (c == drop[0] || c == drop[1] || c == drop[2] || c == drop[3]) continue;
        const double v = fdel_s[c][tid];
        const bool better = bi < 0 || (pass < 2 ? dv_precedes(v, c, bv, bi) : dv_precedes(bv, bi, v, c));
        if (better) {
          bv = v;
          bi = c;
        }
      }
      if (pass == 0) drop[0] = bi; else if (pass == 1) drop[1] = bi; else if (pass == 2) drop[2] = bi; else drop[3] = bi;
    }
    double best_err = INFINITY, best_val = 0.0;
    int best_c = -1;
#pragma unroll 1
    for (int c = 0; c < DV_NEST; ++c) {
      if (c == drop[0] || c == drop[1] || c == drop[2] || c == drop[3]) continue;
      const double v = fdel_s[c][tid], er = err_s[c][tid];
      // min() returns the first minimum in sorted order -> on ties the extrapolant that precedes the other
      if (best_c < 0 || er < best_err || (er == best_err && dv_precedes(v, c, best_val, best_c))) {
        best_err = er;
        best_val = v;
        best_c = c;
      }
    }
    a.grad[o_s[pr] * NJ + s] = best_val;
    if (touched && a.flags) {  // the flag belongs to the item's problem, not to this thread's phase-1 pair
      const long long rest = tp / a.nobs;
      const int slot = (int)(rest / a.H);
      atomicOr(&a.flags[a.list ? a.list[slot] : slot], 0x100);
      touched = 0;
    }
  }
  if (touched && a.flags && prob >= 0) atomicOr(&a.flags[prob], 0x100);
}

template <int NJ>
static cudaError_t launch_derivest_nj(const GradArgs &a, long long total, cudaStream_t s) {
  // 128 pairs per CTA unless the launch would leave SMs without a CTA (late outer iterations): then 32 pairs per CTA, i.e.
  // at most two rounds of work items instead of up to NJ (measured: 51 200 pairs 0.46 ms at 128, 0.50 ms at 32)
  const int ppc = total >= 148LL * GRAD_THREADS ? GRAD_THREADS : 32;
  const int grid = (int)((total + ppc - 1) / ppc);
  const size_t dyn = sizeof(double) * (DV_NDEL + DV_NEST + 3 * NJ) * GRAD_THREADS;  // 60 KB at NJ = 5
  cudaError_t e = cudaFuncSetAttribute(k_grad_derivest<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return e;
  k_grad_derivest<NJ><<<grid, GRAD_THREADS, dyn, s>>>(a, ppc);
  return cudaGetLastError();
}

cudaError_t launch_grad_derivest(const GradArgs &a, cudaStream_t s) {
  const long long total = (long long)a.nslots * a.H * a.nobs;  // (waypoint, obstacle) pairs
  if (total <= 0) return cudaSuccess;
  switch (a.nj) {
    case 2: return launch_derivest_nj<2>(a, total, s);
    case 3: return launch_derivest_nj<3>(a, total, s);
    case 4: return launch_derivest_nj<4>(a, total, s);
    case 5: return launch_derivest_nj<5>(a, total, s);
    case 6: return launch_derivest_nj<6>(a, total, s);
    default: return cudaErrorInvalidValue;
  }
}

// ============================================================================================================
// K6: RRT node feasibility (RRT_FANUC.m:146-181) and nearest/steer (RRT_FANUC.m:116-129)
// ============================================================================================================
__global__ void __launch_bounds__(GRAD_THREADS) k_nodes_feasible(const DevTables *gtab, int nj, int nobs, int N,
                                                                const double *theta, unsigned char *feasible,
                                                                double *dmin) {
  __shared__ alignas(128) DevTables tab;
  __shared__ alignas(8) uint64_t mbar;
  tma_stage(&tab, gtab, tab_bytes(nobs), &mbar);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  Xf M, Mn;
  double p[6];
  int touched = 0;
  bool feas = true;
  double dm = INFINITY;
  for (int l = 0; l < nj; ++l) {
    double s, c;
    sincos(theta[(long long)t * nj + l] + tab.link[l].th_off, &s, &c);
    if (l == 0) {
      xf_first(tab.link[0], c, s, M);
    } else {
      xf_step(M, tab.link[l], c, s, Mn);
      M = Mn;
    }
    link_endpoints(M, tab.link[l], tab.base, p);
    for (int j = 0; j < nobs; ++j) {
      const double d = link_obs_dist(p, tab.obs[j], touched);
      dm = d < dm ? d : dm;
      if (d < tab.obs[j].D) feas = false;  // RRT_FANUC.m:172
    }
  }
  feasible[t] = feas ? 1 : 0;
  if (dmin) dmin[t] = dm;
}

// ---- work order of the fused solver: longest expected problems first ------------------------------------------------------
// A problem whose reference line passes inside an obstacle margin needs several CFS iterations (or a long infeasibility
// certificate); one that stays clear converges in two.  count[b] = number of (waypoint, obstacle) pairs of x_ with
// distance < margin; k_order_desc then lists the problems by descending count (counting sort, one CTA), and the persistent
// CTAs of k_cfs_fused pull them in that order: the long problems start first instead of wherever the batch put them.
// The order only changes WHEN a problem is solved, never its result (every problem is solved independently).
__global__ void __launch_bounds__(GRAD_THREADS) k_difficulty(const DevTables *gtab, int nj, int nobs, int B, int H,
                                                            const double *xref, int margin_is_D, int *count) {
  __shared__ alignas(128) DevTables tab;
  __shared__ alignas(8) uint64_t mbar;
  tma_stage(&tab, gtab, tab_bytes(nobs), &mbar);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * H) return;
  const int b = (int)(t / H);
  const double *th = xref + t * 2 * nj;
  Xf M;
  double p[6];
  int touched = 0, viol = 0;
  double dm[CFS_MAX_OBS];
  for (int j = 0; j < nobs; ++j) dm[j] = INFINITY;
  for (int l = 0; l < nj; ++l) {
    double sn, cs;
    sincos(th[l] + tab.link[l].th_off, &sn, &cs);
    if (l == 0)
      xf_first(tab.link[0], cs, sn, M);
    else
      xf_step_inplace(M, tab.link[l], cs, sn);
    link_endpoints(M, tab.link[l], tab.base, p);
    for (int j = 0; j < nobs; ++j) {
      const double k = link_obs_key(p, tab.obs[j], touched);
      dm[j] = k < dm[j] ? k : dm[j];
    }
  }
  for (int j = 0; j < nobs; ++j) viol += key_to_dist(dm[j]) < (margin_is_D ? tab.obs[j].D : tab.obs[j].eps) ? 1 : 0;
  if (viol) atomicAdd(&count[b], viol);
}

__global__ void __launch_bounds__(1024) k_order_desc(int B, int max_count, const int *count, int *order) {
  extern __shared__ int hist[];  // max_count + 2
  for (int e = threadIdx.x; e <= max_count + 1; e += blockDim.x) hist[e] = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) atomicAdd(&hist[min(count[b], max_count)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {  // start offset of every bucket, largest count first
    int o = 0;
    for (int c = max_count; c >= 0; --c) {
      const int h = hist[c];
      hist[c] = o;
      o += h;
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) order[atomicAdd(&hist[min(count[b], max_count)], 1)] = b;
}

cudaError_t launch_work_order(const DevTables *tab, int nj, int nobs, int B, int H, const double *xref, int margin_is_D,
                              int *count, int *order, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int) * B, s);
  if (e != cudaSuccess) return e;
  const long long total = (long long)B * H;
  k_difficulty<<<(int)((total + GRAD_THREADS - 1) / GRAD_THREADS), GRAD_THREADS, 0, s>>>(tab, nj, nobs, B, H, xref, margin_is_D,
                                                                                       count);
  const int max_count = H * nobs;
  k_order_desc<<<1, 1024, sizeof(int) * (max_count + 2), s>>>(B, max_count, count, order);
  return cudaGetLastError();
}

cudaError_t launch_nodes_feasible(const DevTables *tab, int nj, int nobs, int N, const double *theta,
                                  unsigned char *feasible, double *dmin, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  k_nodes_feasible<<<(N + GRAD_THREADS - 1) / GRAD_THREADS, GRAD_THREADS, 0, s>>>(tab, nj, nobs, N, theta, feasible,
                                                                                   dmin);
  return cudaGetLastError();
}

__global__ void k_nearest_steer(int nj, int n_nodes, const double *nodes, int S, const double *samples,
                                const double *ratial, double step, int *parent, double *newnode) {
  const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= S) return;
  double smp[CFS_MAXL], rat[CFS_MAXL];
  for (int k = 0; k < nj; ++k) {
    smp[k] = samples[(long long)sidx * nj + k];
    rat[k] = ratial[k];
  }
  int best = 0;
  double bestd = 0.0;
  for (int i = 0; i < n_nodes; ++i) {
    double ss = 0.0;
    for (int k = 0; k < nj; ++k) {
      const double v = (nodes[(long long)i * nj + k] - smp[k]) * rat[k];
      ss += v * v;
    }
    const double d = sqrt(ss);
    if (i == 0 || d < bestd) {  // RRT_FANUC.m:119-126 : strict <, first minimum wins
      bestd = d;
      best = i;
    }
  }
  parent[sidx] = best;
  double ss = 0.0;
  for (int k = 0; k < nj; ++k) {
    const double v = nodes[(long long)best * nj + k] - smp[k];
    ss += v * v;
  }
  const double nrm = sqrt(ss);
  for (int k = 0; k < nj; ++k) {  // RRT_FANUC.m:129
    const double pk = nodes[(long long)best * nj + k];
    newnode[(long long)sidx * nj + k] = pk + (smp[k] - pk) * step / nrm;
  }
}

cudaError_t launch_nearest_steer(int nj, int n_nodes, const double *nodes, int S, const double *samples,
                                 const double *ratial, double step, int *parent, double *newnode, cudaStream_t s) {
  if (S <= 0) return cudaSuccess;
  k_nearest_steer<<<(S + 127) / 128, 128, 0, s>>>(nj, n_nodes, nodes, S, samples, ratial, step, parent, newnode);
  return cudaGetLastError();
}

// ============================================================================================================
// FP64 FMA peak micro-benchmark (roofline denominator)
// ============================================================================================================
__global__ void k_fp64_peak(double *sink, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 123.456) sink[0] = r;
}

cudaError_t launch_fp64_peak(double *sink, int iters, int grid, int block, cudaStream_t s) {
  k_fp64_peak<<<grid, block, 0, s>>>(sink, iters);
  return cudaGetLastError();
}

}  // namespace cfs
