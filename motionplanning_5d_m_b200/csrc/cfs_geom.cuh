// cfs_geom.cuh -- device-side forward kinematics and capsule-axis distances, FP64, all in registers.
//
// Replaces (reference file:line):
//   Lib/functions/CapPos.m:8-23          chain  M{i+1}=M{i}*[R T;0 0 0 1], endpoints pos{i}.p
//   Lib/2L/CapPos2.m:16-29               planar chain with constant translations robot.T(:,i)
//   Lib/functions/distLinSeg.m:23-101    Lumelsky segment-segment distance
//   Lib/M16iB/dist_arm_3D_Heu_2.m:20-29, Lib/200i/dist_arm_3D_200i_2.m:21-29, Lib/2L/dist_arm_2L.m:13-22
//                                        min over links + "negative when the axes touch" rule
#pragma once
#include "cfs_types.cuh"

namespace cfs {

// ---- TMA bulk staging of the robot/obstacle tables -----------------------------------------------------
// One elected thread arms an mbarrier with the byte count and issues cp.async.bulk (SASS: UBLKCP);
// every thread then waits on the barrier phase.  The tables are < 4 KB, so this is about latency, not
// bandwidth: one bulk transaction instead of ~100 scattered LDGs per thread.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void tma_stage(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *mbar) {
  const uint32_t bar = smem_u32(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(bar)
        : "memory");
  }
  // phase 0 wait
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(0u)
        : "memory");
  }
}

__device__ __forceinline__ uint32_t tab_bytes(int nobs) {
  return static_cast<uint32_t>(CFS_TAB_HEADER_BYTES + sizeof(ObsTab) * nobs);
}

// ---- one chain step: Mn = M * [R T; 0 0 0 1]   (CapPos.m:13-17) ------------------------------------------
// M is 3x4 row-major in registers (bottom row is [0 0 0 1] throughout).  R(3,1) == 0 structurally.
struct Xf {
  double m[12];
};

__device__ __forceinline__ void link_RT(const LinkTab &L, double c, double s, double R[9], double T[3]) {
  R[0] = c;   R[1] = -s * L.ca;  R[2] = s * L.sa;
  R[3] = s;   R[4] = c * L.ca;   R[5] = -c * L.sa;
  R[6] = 0.0; R[7] = L.sa;       R[8] = L.ca;
  T[0] = L.a * c + L.tx;
  T[1] = L.a * s + L.ty;
  T[2] = L.dz;
}

__device__ __forceinline__ void xf_first(const LinkTab &L, double c, double s, Xf &M) {
  double R[9], T[3];
  link_RT(L, c, s, R, T);  // eye(4) * X == X exactly
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    M.m[4 * a + 0] = R[3 * a + 0];
    M.m[4 * a + 1] = R[3 * a + 1];
    M.m[4 * a + 2] = R[3 * a + 2];
    M.m[4 * a + 3] = T[a];
  }
}

__device__ __forceinline__ void xf_step(const Xf &M, const LinkTab &L, double c, double s, Xf &Mn) {
  double R[9], T[3];
  link_RT(L, c, s, R, T);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double m0 = M.m[4 * a], m1 = M.m[4 * a + 1], m2 = M.m[4 * a + 2], m3 = M.m[4 * a + 3];
    Mn.m[4 * a + 0] = m0 * R[0] + m1 * R[3];  // + m2*0
    Mn.m[4 * a + 1] = (m0 * R[1] + m1 * R[4]) + m2 * R[7];
    Mn.m[4 * a + 2] = (m0 * R[2] + m1 * R[5]) + m2 * R[8];
    Mn.m[4 * a + 3] = ((m0 * T[0] + m1 * T[1]) + m2 * T[2]) + m3;
  }
}

// same products as xf_step, M updated in place (row a of the product only reads row a of M)
__device__ __forceinline__ void xf_step_inplace(Xf &M, const LinkTab &L, double c, double s) {
  double R[9], T[3];
  link_RT(L, c, s, R, T);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double m0 = M.m[4 * a], m1 = M.m[4 * a + 1], m2 = M.m[4 * a + 2], m3 = M.m[4 * a + 3];
    M.m[4 * a + 0] = m0 * R[0] + m1 * R[3];  // + m2*0
    M.m[4 * a + 1] = (m0 * R[1] + m1 * R[4]) + m2 * R[7];
    M.m[4 * a + 2] = (m0 * R[2] + m1 * R[5]) + m2 * R[8];
    M.m[4 * a + 3] = ((m0 * T[0] + m1 * T[1]) + m2 * T[2]) + m3;
  }
}

// M <- M * [R T; 0 0 0 1] with R, T already built by link_RT: the same products as xf_step_inplace (two chains that take the
// same link step -- same joint angle -- share R and T)
__device__ __forceinline__ void xf_apply_inplace(Xf &M, const double R[9], const double T[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double m0 = M.m[4 * a], m1 = M.m[4 * a + 1], m2 = M.m[4 * a + 2], m3 = M.m[4 * a + 3];
    M.m[4 * a + 0] = m0 * R[0] + m1 * R[3];  // + m2*0
    M.m[4 * a + 1] = (m0 * R[1] + m1 * R[4]) + m2 * R[7];
    M.m[4 * a + 2] = (m0 * R[2] + m1 * R[5]) + m2 * R[8];
    M.m[4 * a + 3] = ((m0 * T[0] + m1 * T[1]) + m2 * T[2]) + m3;
  }
}

// pos{i}.p(:,k) = M(1:3,1:3)*cap.p(:,k) + M(1:3,4) + base   (CapPos.m:18-20); p[0..2]=start, p[3..5]=end
__device__ __forceinline__ void link_endpoints(const Xf &M, const LinkTab &L, const double base[3], double p[6]) {
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a)
      p[3 * k + a] = (((M.m[4 * a] * L.cap[k][0] + M.m[4 * a + 1] * L.cap[k][1]) + M.m[4 * a + 2] * L.cap[k][2]) +
                      M.m[4 * a + 3]) + base[a];
}

// x / D2 for the obstacle constant D2 with its correctly rounded reciprocal r: q0 = RN(x r), rem = x - q0 D2 (exact, FMA),
// q = RN(q0 + rem r) is the correctly rounded quotient (Markstein's theorem; x, D2 are far from over/underflow here)
__device__ __forceinline__ double div_by_const(double x, double d, double r) {
  const double q0 = x * r;
  const double rem = fma(-q0, d, x);
  return fma(rem, r, q0);
}

__device__ __forceinline__ double fixbound(double v) {  // distLinSeg.m:93-101
  return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
}

// distLinSeg(link start, link end, obs start, obs end) followed by the touch rule of dist_arm_*:
//   if |dis| < 1e-4: dis = -norm(P1 - link_end)       (dist_arm_3D_200i_2.m:22-24, dist_link_Heu.m:19-21)
// The branch-deciding quantity den = D1*D2 - R^2 is formed with explicit round-to-nearest mul/sub (no FMA
// contraction) so that the parallel-lines test (distLinSeg.m:55) sees the same value as MATLAB's arithmetic.
//
// link_obs_key returns the signed square of that distance: key = |v|^2 for an ordinary link, -|w|^2 for a touching one.
// sqrt is monotone and correctly rounded, so min over links of the distances == key_to_dist(min over links of the keys)
// bit for bit, and one evaluation of dist_arm needs ONE square root instead of one per link (35 -> 11 per num_jac
// waypoint).  CFS_TOUCH_KEY = 1e-08 (0x1.5798ee2308c3ap-27) is the smallest double whose square root is >= 1e-4, so
// "key < CFS_TOUCH_KEY" is exactly the reference's "abs(dis) < 1e-4".  (Only difference: when two links' distances round
// to the same double from different squares, the first-minimal-link rule may name the other link; the value is the same.)
#define CFS_TOUCH_KEY 1e-08
__device__ __forceinline__ double key_to_dist(double k) { return k >= 0.0 ? sqrt(k) : -sqrt(-k); }

// ---- N3 extension: solid axis-aligned box obstacles (obs{j}.shape = 'box', l = [min corner, max corner]) ----------------------
// Squared distance between the link axis [p[0..2], p[3..5]] and the box, with the same touch rule as the capsule obstacles.
// With x(t) = ps + t d and the signed per-axis excess e_k(x) (x_k - hi_k above, x_k - lo_k below, 0 inside), f(t) = sum e_k^2 is
// convex and C1: g(t) = f'/2 = sum_k e_k d_k is piecewise linear and nondecreasing, its breakpoints are the (<= 6) parameters
// where x_k crosses lo_k / hi_k.  t* = 0 if g(0) >= 0, 1 if g(1) <= 0, else the root of g: the nearest breakpoints on either
// side of it are kept while g is evaluated at every breakpoint inside (0,1), then one linear interpolation.  Same algorithm,
// same operation order as orc_dist_seg_box (oracle/cfs_oracle.c).
__device__ __forceinline__ double box_excess_dot(const double p[6], const double d[3], const ObsTab &o, double t, double &ss) {
  double g = 0.0;
  ss = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double x = p[k] + d[k] * t;
    const double e = x > o.d2[k] ? x - o.d2[k] : (x < o.s[k] ? x - o.s[k] : 0.0);
    g += e * d[k];
    ss += e * e;
  }
  return g;
}

// Not inlined (rare path, keeps the capsule path's code and registers as they were); scalar arguments only, so nothing goes
// through local memory.  A negative key <=> the touch branch was taken.
static __device__ __noinline__ double link_box_key(double p0, double p1, double p2, double p3, double p4, double p5, const ObsTab *op) {
  const ObsTab &o = *op;
  const double p[6] = {p0, p1, p2, p3, p4, p5};
  double d[3], ss;
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = p[3 + k] - p[k];
  double t;
  const double g0 = box_excess_dot(p, d, o, 0.0, ss);
  if (g0 >= 0.0) {
    t = 0.0;
  } else {
    const double g1 = box_excess_dot(p, d, o, 1.0, ss);
    if (g1 < 0.0) {
      t = 1.0;
    } else {
      double tl = 0.0, gl = g0, th = 1.0, gh = g1;
#pragma unroll 1
      for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.0) continue;
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
          const double c = ((side ? o.d2[k] : o.s[k]) - p[k]) / d[k];
          if (!(c > 0.0 && c < 1.0)) continue;
          const double gc = box_excess_dot(p, d, o, c, ss);
          if (gc < 0.0) {
            if (c > tl) { tl = c; gl = gc; }
          } else if (c < th) {
            th = c; gh = gc;
          }
        }
      }
      t = tl - gl * (th - tl) / (gh - gl);  // gl < 0 <= gh; gh == 0 gives t = th, the FIRST minimiser (entry point of a crossing)
    }
  }
  (void)box_excess_dot(p, d, o, t, ss);
  double key = ss;
  if (key < CFS_TOUCH_KEY) {  // |dis| < 1e-4: dis = -norm(P1 - link end), the rule of dist_arm_3D_200i_2.m:22-24 applied to the box
    const double wx = (p[0] + d[0] * t) - p[3], wy = (p[1] + d[1] * t) - p[4], wz = (p[2] + d[2] * t) - p[5];
    key = -((wx * wx + wy * wy) + wz * wz);
  }
  return key;
}

__device__ __forceinline__ double link_obs_key(const double p[6], const ObsTab &o, int &touched) {
  if (o.kind == CFS_OBS_BOX) {
    const double key = link_box_key(p[0], p[1], p[2], p[3], p[4], p[5], &o);
    if (key < 0.0) touched = 1;
    return key;
  }
  const double d1x = p[3] - p[0], d1y = p[4] - p[1], d1z = p[5] - p[2];
  const double d12x = o.s[0] - p[0], d12y = o.s[1] - p[1], d12z = o.s[2] - p[2];
  const double d2x = o.d2[0], d2y = o.d2[1], d2z = o.d2[2];
  const double D1 = (d1x * d1x + d1y * d1y) + d1z * d1z;
  const double D2 = o.D2;
  const double S1 = (d1x * d12x + d1y * d12y) + d1z * d12z;
  const double S2 = (d2x * d12x + d2y * d12y) + d2z * d12z;
  const double R = (d1x * d2x + d1y * d2y) + d1z * d2z;
  const double den = __dsub_rn(__dmul_rn(D1, D2), __dmul_rn(R, R));
  double t, u;
  if (D1 == 0.0 || D2 == 0.0) {
    if (D1 != 0.0) {
      u = 0.0;
      t = fixbound(S1 / D1);
    } else if (D2 != 0.0) {
      t = 0.0;
      u = fixbound(div_by_const(-S2, D2, o.rD2));
    } else {
      t = 0.0;
      u = 0.0;
    }
  } else if (den == 0.0) {
    t = 0.0;
    u = div_by_const(-S2, D2, o.rD2);
    const double uf = fixbound(u);
    if (uf != u) {
      t = fixbound((uf * R + S1) / D1);
      u = uf;
    }
  } else {
    t = fixbound((S1 * D2 - S2 * R) / den);
    u = div_by_const(t * R - S2, D2, o.rD2);
    const double uf = fixbound(u);
    if (uf != u) {
      t = fixbound((uf * R + S1) / D1);
      u = uf;
    }
  }
  const double vx = (d1x * t - d2x * u) - d12x, vy = (d1y * t - d2y * u) - d12y, vz = (d1z * t - d2z * u) - d12z;
  double key = (vx * vx + vy * vy) + vz * vz;
  if (key < CFS_TOUCH_KEY) {
    const double wx = (p[0] + d1x * t) - p[3], wy = (p[1] + d1y * t) - p[4], wz = (p[2] + d1z * t) - p[5];
    key = -((wx * wx + wy * wy) + wz * wz);
    touched = 1;
  }
  return key;
}

__device__ __forceinline__ double link_obs_dist(const double p[6], const ObsTab &o, int &touched) {
  return key_to_dist(link_obs_key(p, o, touched));
}

}  // namespace cfs
