// k_rrt.cu -- batched RRT / RRT* tree growth: one CTA per seed, sm_100a.
//
// Replaces RRT_FANUC.find_route (Lib/RRT_FANUC.m:63-93) with getNode / getRandNode (:97-131), feasible (:146-181),
// addNode (:184-190), arrangeNode (:134-142) and goal_reached (:193-207) for S independent seeds -- the GPU analogue of the
// parfor over num_seed workers in Lib/functions/s_Parallel_rrt.m:16-25 (routeL(i) = size(route,2), then min over seeds).
//
// A tree is sequential by construction (every sample is steered from the nearest node of the tree so far), so the
// parallelism is across seeds (one CTA each, the whole tree -- <= MAX_ITER+1 nodes, parents, path lengths, sample
// distances -- in shared memory) and, inside a seed, across the nodes of the nearest-neighbour scan and the RRT* re-parenting
// pass and across the obstacles of the feasibility test.  MATLAB's rand stream is an input (rnd, consumed in the reference's
// order: pp = rand, then rand(nstate,1) when pp < bi), so a run is reproducible against the CPU restatement; the arithmetic
// that decides the tree (distances, steering, path lengths, goal box) is written with explicit round-to-nearest operations
// in the reference's order (no FMA contraction).
#include "cfs_geom.cuh"
#include "cfs_kernels.cuh"

namespace cfs {

#define RRT_THREADS 128

struct RrtLayout {
  size_t nodes, total, todis, parent, total_bytes;
};
__host__ __device__ inline RrtLayout rrt_layout(int nj, int cap) {
  RrtLayout L;
  size_t o = 0;
  L.nodes = o; o += sizeof(double) * (size_t)nj * cap;
  L.total = o; o += sizeof(double) * cap;
  L.todis = o; o += sizeof(double) * cap;
  L.parent = o; o += sizeof(int) * cap;
  L.total_bytes = (o + 15) / 16 * 16;
  return L;
}

// first strict minimum over the CTA (ties: lower index), every thread returns the same pair
__device__ __forceinline__ void rrt_argmin(double &val, int &idx, double *red_v, int *red_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, val, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (oi >= 0 && (idx < 0 || ov < val || (ov == val && oi < idx))) { val = ov; idx = oi; }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { red_v[w] = val; red_i[w] = idx; }
  __syncthreads();
  val = red_v[0]; idx = red_i[0];
#pragma unroll
  for (int k = 1; k < RRT_THREADS / 32; ++k) {
    const double ov = red_v[k];
    const int oi = red_i[k];
    if (oi >= 0 && (idx < 0 || ov < val || (ov == val && oi < idx))) { val = ov; idx = oi; }
  }
}

__global__ void __launch_bounds__(RRT_THREADS) k_rrt_find_routes(RrtArgs a) {
  extern __shared__ __align__(16) unsigned char rrt_smem[];
  __shared__ alignas(128) DevTables tab;
  __shared__ alignas(8) uint64_t mbar;
  __shared__ double s_sample[CFS_MAXL], s_new[CFS_MAXL], red_v[RRT_THREADS / 32], s_cs[2 * CFS_MAXL], s_p[6 * CFS_MAXL];
  __shared__ int red_i[RRT_THREADS / 32], s_ctl[4];
  __shared__ double s_rnd[2 * RRT_THREADS];  // window of the seed's random stream (refilled by the whole CTA)
  const int nj = a.nj, cap = a.max_iter + 2, tid = threadIdx.x, seed = blockIdx.x;
  const RrtLayout L = rrt_layout(nj, cap);
  double *nodes = reinterpret_cast<double *>(rrt_smem + L.nodes);
  double *total = reinterpret_cast<double *>(rrt_smem + L.total);
  double *todis = reinterpret_cast<double *>(rrt_smem + L.todis);
  int *parent = reinterpret_cast<int *>(rrt_smem + L.parent);
  tma_stage(&tab, a.tab, tab_bytes(a.nobs), &mbar);
  const double *x0 = a.x0 + (size_t)seed * nj, *goal = a.goal + (size_t)seed * nj, *goal_th = a.goal_th + (size_t)seed * nj;
  const double *rnd = a.rnd + (size_t)seed * a.nrnd;
  if (tid < nj) { nodes[tid] = x0[tid]; s_new[tid] = x0[tid]; }
  if (tid == 0) { parent[0] = -1; total[0] = 0.0; }
  int node_num = 1, cur = 0, par = 0, fail = 0, exhausted = 0, touched = 0, win = -(1 << 30);
  __syncthreads();
  for (;;) {
    // ---- goal_reached (:193-207): every joint inside the goal box ----
    int in = 1;
    for (int k = 0; k < nj; ++k)
      if (!(__dsub_rn(goal[k], a.region_g[k]) < s_new[k] && s_new[k] < __dadd_rn(goal[k], a.region_g[k]))) in = 0;
    if (node_num > a.max_iter) { fail = 1; in = 1; }
    if (in) break;
    // ---- getNode (:97-104): sample, nearest, steer, until feasible ----
    for (;;) {
      __syncthreads();
      if (cur + nj + 1 > win + 2 * RRT_THREADS) {  // the next draw (1 + nj numbers) must lie inside the window
        win = cur;
        for (int e = tid; e < 2 * RRT_THREADS; e += RRT_THREADS) s_rnd[e] = win + e < a.nrnd ? rnd[win + e] : 0.0;
        __syncthreads();
      }
      if (tid == 0) {
        int c = cur, ex = 0;
        if (c >= a.nrnd) {
          ex = 1;
        } else {
          const double pp = s_rnd[c++ - win];
          if (pp < a.bi) {
            if (c + nj > a.nrnd) {
              ex = 1;
            } else {
              for (int k = 0; k < nj; ++k)
                s_sample[k] = __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(s_rnd[c + k - win], 0.5), a.region_s[k]), 2.0), a.sample_off[k]);
              c += nj;
            }
          } else {
            for (int k = 0; k < nj; ++k) s_sample[k] = goal_th[k];
          }
        }
        s_ctl[0] = c;
        s_ctl[1] = ex;
      }
      __syncthreads();
      cur = s_ctl[0];
      if (s_ctl[1]) { exhausted = 1; break; }
      // nearest (:116-127): weighted distance of the sample to every node of the tree
      double best = 0.0;
      int bidx = -1;
      for (int i = tid; i < node_num; i += RRT_THREADS) {
        double ss = 0.0;
        for (int k = 0; k < nj; ++k) {
          const double v = __dmul_rn(__dsub_rn(nodes[i * nj + k], s_sample[k]), a.ratial[k]);
          ss = __dadd_rn(ss, __dmul_rn(v, v));
        }
        const double d = sqrt(ss);
        todis[i] = d;
        if (bidx < 0 || d < best) { best = d; bidx = i; }
      }
      rrt_argmin(best, bidx, red_v, red_i);
      par = bidx + 1;
      // steer (:129): parent + (sample - parent) * 0.1 / norm(parent - sample)
      if (tid == 0) {
        const double *pn = nodes + bidx * nj;
        double ss = 0.0;
        for (int k = 0; k < nj; ++k) {
          const double v = __dsub_rn(pn[k], s_sample[k]);
          ss = __dadd_rn(ss, __dmul_rn(v, v));
        }
        const double nrm = sqrt(ss);
        for (int k = 0; k < nj; ++k) s_new[k] = __dadd_rn(pn[k], __ddiv_rn(__dmul_rn(__dsub_rn(s_sample[k], pn[k]), 0.1), nrm));
      }
      __syncthreads();
      // feasible (:146-181): sin/cos per joint lane, the chain of link transforms on one thread (the only sequential part),
      // then one thread per (link, obstacle) pair; infeasible if any link is closer than obs{j}.D
      if (tid < nj) {
        double sn, cs;
        sincos(s_new[tid] + tab.link[tid].th_off, &sn, &cs);
        s_cs[tid] = cs;
        s_cs[CFS_MAXL + tid] = sn;
      }
      __syncthreads();
      if (tid == 0) {
        Xf M;
        for (int l = 0; l < nj; ++l) {
          if (l == 0)
            xf_first(tab.link[0], s_cs[0], s_cs[CFS_MAXL], M);
          else
            xf_step_inplace(M, tab.link[l], s_cs[l], s_cs[CFS_MAXL + l]);
          link_endpoints(M, tab.link[l], tab.base, &s_p[6 * l]);
        }
      }
      __syncthreads();
      int ok = 1;
      for (int e = tid; e < nj * a.nobs; e += RRT_THREADS) {
        const int l = e % nj, j = e / nj;
        if (link_obs_dist(&s_p[6 * l], tab.obs[j], touched) < tab.obs[j].D) ok = 0;
      }
      if (__syncthreads_and(ok)) break;
    }
    if (exhausted) break;
    // ---- addNode (:184-190) ----
    if (tid < nj) nodes[node_num * nj + tid] = s_new[tid];
    if (tid == 0) {
      parent[node_num] = par;
      total[node_num] = __dadd_rn(total[par - 1], todis[par - 1]);
    }
    ++node_num;
    __syncthreads();
    // ---- arrangeNode (:134-142): nodes closer than 0.2 to the SAMPLE re-parent to the new node when that is shorter ----
    if (a.star) {
      const double tn = total[node_num - 1];
      for (int i = tid; i < node_num - 1; i += RRT_THREADS) {
        const double cand = __dadd_rn(tn, todis[i]);
        if (todis[i] < 0.2 && total[i] > cand) {
          parent[i] = node_num;
          total[i] = cand;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- route (:86-91): walk the parents back from the last node ----
  if (tid == 0) {
    int len = 1, p = par, guard = 0;
    if (exhausted) {
      len = -1;
    } else {
      while (p > 0 && guard++ <= node_num) { ++len; p = parent[p - 1]; }
      double *route = a.routes + (size_t)seed * nj * (a.max_iter + 2);
      int pos = len - 1;
      for (int k = 0; k < nj; ++k) route[(size_t)pos * nj + k] = s_new[k];
      p = par; guard = 0;
      while (p > 0 && guard++ <= node_num) {
        --pos;
        for (int k = 0; k < nj; ++k) route[(size_t)pos * nj + k] = nodes[(size_t)(p - 1) * nj + k];
        p = parent[p - 1];
      }
    }
    a.route_len[seed] = len;
    a.n_nodes[seed] = node_num;
    a.fail[seed] = fail;
    a.rnd_used[seed] = cur;
  }
  if (a.tree_nodes) {  // optional dump of the whole tree (tests)
    for (int e = tid; e < node_num * nj; e += RRT_THREADS) a.tree_nodes[(size_t)seed * nj * cap + e] = nodes[e];
    for (int e = tid; e < node_num; e += RRT_THREADS) {
      a.tree_parent[(size_t)seed * cap + e] = parent[e];
      a.tree_total[(size_t)seed * cap + e] = total[e];
    }
  }
  (void)touched;
}

cudaError_t launch_rrt_find_routes(const RrtArgs &a, int S, cudaStream_t s) {
  if (S <= 0) return cudaSuccess;
  const size_t smem = rrt_layout(a.nj, a.max_iter + 2).total_bytes;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_rrt_find_routes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_rrt_find_routes<<<S, RRT_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace cfs
