// k_chomp.cu -- CHOMP_FANUC (Lib/CHOMP_FANUC.m), the gradient-descent baseline planner, batched.
//
// One outer iteration of CHOMP_FANUC.optimizer (:56-68) for B problems is
//     K1d in dm_f mode (k_grad.cu)   dm_f (:115-134) per link, linkid = argmin, derivest(dist_link_*(linkid)) per joint (:151,:156)
//     k_chomp_step                   dcostObs_f (:137-165) chained through Baug, u <- u - alpha*3*(QQ u + ff + 2000 dcostObs) (:75),
//                                    roll-out (:77-82), e_u = ||u_old - u|| (EVAL.m:58)
//     k_dgemm (k_setup.cu)           QQ u for the whole batch (cost now, dcostArm_f of the next iteration: :87)
//     k_chomp_cost                   get_cost(u) + fobs_m() (:63, :91-112)
// The stop rule never fires before MAX_O_ITER (eval.x_ / eval.x_old are never updated by CHOMP_FANUC), so every problem
// runs exactly max_outer iterations: no active list, no stragglers -- plain data-parallel launches.
//
// Faithful quirk (:153,:158): the gradient of waypoint i is chained through Baug((i-1)*njoint+1 : i*njoint, :) -- a row stride
// of njoint where the state blocks of Baug have nstate = 2*njoint rows -- i.e. the THETA rows of step (i+1)/2 for odd i and
// the OMEGA rows of step i/2 for even i (1-based).  Restated as written; oracle: orc_chomp_solve.
#include "cfs_kernels.cuh"

namespace cfs {

#define CHOMP_THREADS 128

template <int NT>
__device__ __forceinline__ double chomp_block_sum(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) s += red[w];  // fixed order
  return s;
}

// shared memory: wk[OH*nj] (w_ij * dDfx_ijk) | red[8]
__global__ void __launch_bounds__(CHOMP_THREADS) k_chomp_step(ChompArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, n = a.n, nj = a.nj, H = a.H, OH = a.nobs * H;
  double *wk = sm, *red = sm + (size_t)OH * nj;
  const double dt = a.tab->dt;
  // weight of every (obstacle, waypoint) pair times the derivest gradient of its closest link   (:147-161)
  for (int e = tid; e < OH; e += CHOMP_THREADS) {
    const int j = e / H;
    const long long o = (long long)b * OH + e;
    const double d = a.dist[o] - a.tab->obs[j].D, eps = a.tab->obs[j].eps;
    double w = 0.0;
    if (d < 0.0) w = -1.0;
    else if (d <= eps) w = (1.0 / eps) * (d - eps);
    for (int k = 0; k < nj; ++k) wk[e * nj + k] = w != 0.0 ? __dmul_rn(w, a.grad[o * nj + k]) : 0.0;
  }
  __syncthreads();
  // one thread per control (jj, k): the pairs in the reference's order (waypoints outer, obstacles inner)
  double pe = 0.0;
  for (int c = tid; c < n; c += CHOMP_THREADS) {
    const int jj = c / nj + 1, k = c - (jj - 1) * nj;
    double g = 0.0;
    for (int i = 1; i <= H; ++i) {
      const int step = (i + 1) / 2;
      if (step < jj) continue;
      const double bcoef = (i & 1) ? __dadd_rn(0.5 * dt * dt, __dmul_rn((step - jj) * dt, dt)) : dt;
      for (int j = 0; j < a.nobs; ++j) {
        const double v = wk[(j * H + (i - 1)) * nj + k];
        if (v != 0.0) g = __dadd_rn(g, __dmul_rn(v, bcoef));
      }
    }
    const size_t gi = (size_t)b * n + c;
    const double uo = a.u[gi];
    const double un = __dsub_rn(uo, __dmul_rn(a.alpha * 3, __dadd_rn(__dadd_rn(a.w[gi], a.ff[gi]), __dmul_rn(2000.0, g))));
    a.u[gi] = un;
    pe += (uo - un) * (uo - un);
  }
  const double e2 = chomp_block_sum<CHOMP_THREADS>(pe, red);  // (also orders the writes of u before the roll-out)
  if (tid == 0 && a.e_u_hist) a.e_u_hist[(size_t)b * a.max_outer + (a.it - 1)] = sqrt(e2);
  __threadfence_block();
  __syncthreads();
  // roll-out (:77-82), one thread per joint
  if (tid < nj) {
    const double *x0 = a.x0 + (size_t)b * 2 * nj;
    double th = x0[tid], om = x0[nj + tid];
    double *xr = a.x + (size_t)b * 2 * n;
    for (int i = 0; i < H; ++i) {
      const double uk = a.u[(size_t)b * n + i * nj + tid];
      const double thn = (th + dt * om) + (0.5 * dt * dt) * uk;
      const double omn = om + dt * uk;
      th = thn;
      om = omn;
      xr[i * 2 * nj + tid] = th;
      xr[i * 2 * nj + nj + tid] = om;
    }
  }
}

__global__ void __launch_bounds__(CHOMP_THREADS) k_chomp_cost(ChompArgs a) {
  __shared__ double red[CHOMP_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x, n = a.n, nj = a.nj, H = a.H, OH = a.nobs * H;
  double quad = 0.0, lin = 0.0, fobs = 0.0;
  for (int c = tid; c < n; c += CHOMP_THREADS) {
    const size_t gi = (size_t)b * n + c;
    quad += a.w[gi] * a.u[gi];
    lin += a.ff[gi] * a.u[gi];
  }
  for (int e = tid; e < OH * nj; e += CHOMP_THREADS) {  // fobs_m (:95-110): every link of every (waypoint, obstacle) pair
    const int pair = e / nj, j = pair / H;
    const double d = a.linkdist[(long long)b * OH * nj + e] - a.tab->obs[j].D, eps = a.tab->obs[j].eps;
    if (d < 0.0) fobs += -d + 0.5 * eps;
    else if (d <= eps) fobs += (1.0 / (2.0 * eps)) * (d - eps) * (d - eps);
  }
  quad = chomp_block_sum<CHOMP_THREADS>(quad, red);
  lin = chomp_block_sum<CHOMP_THREADS>(lin, red);
  fobs = chomp_block_sum<CHOMP_THREADS>(fobs, red);
  if (tid == 0) a.cost_hist[(size_t)b * a.max_outer + (a.it - 1)] = ((0.5 * quad + lin) + a.caug[b]) + fobs;
}

cudaError_t launch_chomp_step(const ChompArgs &a, cudaStream_t s) {
  const size_t smem = sizeof(double) * ((size_t)a.nobs * a.H * a.nj + 8);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_chomp_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_chomp_step<<<a.B, CHOMP_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_chomp_cost(const ChompArgs &a, cudaStream_t s) {
  k_chomp_cost<<<a.B, CHOMP_THREADS, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace cfs
