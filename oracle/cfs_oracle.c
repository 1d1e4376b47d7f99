/*
 * cfs_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See cfs_oracle.h.
 *
 * A from-scratch C restatement of the reference's MATLAB algorithm; each function cites the
 * reference file:line it follows.  Compile with -ffp-contract=off so that the operation
 * order written here is the operation order executed.
 *
 * Pinning status.  PINNED against the reference's own known answers (tests/test_oracle.py, tests/golden/kat.json):
 * distLinSeg doc example (distLinSeg.m:15-18), derivest(@exp,1) (derivest.m:163-174), the DERIVEST demo values
 * (demo/derivest_demo.m:13,31,71,82), robotproperty2 constants, plus an independent numpy restatement.
 * PARITY UNPINNED for quadprog (Optimization Toolbox, not in the reference tree, no MATLAB/Octave offline): the dense
 * Goldfarb-Idnani solver below stands in for it; the QPs are strictly convex, so the optimum is unique.  No golden
 * OUTPUTS of CFS/PSGCFS exist in the reference; tests/golden/cases.npz are outputs of THIS oracle (regression pins).
 */
#include "cfs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Robot constants.  Lib/functions/robotproperty2.m:12-55 (M200i), :58-99 (M16iB), :102-130 (2L)
 * ---------------------------------------------------------------------------------------- */
static const double PI_ = 3.14159265358979323846;

void orc_robot_init(orc_robot *r, int kind) {
  memset(r, 0, sizeof(*r));
  r->kind = kind;
  r->dt = 0.5; /* robotproperty2.m:17,65,109 */
  if (kind == ORC_M200I) {
    /* robotproperty2.m:24-29 */
    const double DH[6][4] = {{0, 0, 0.050, -1.5708},      {-1.5708, 0, 0.440, 3.1416}, {0, 0, 0.035, -1.5708},
                             {0, -0.420, 0, 1.5708},      {0, 0, 0, -1.5708},          {0, -0.080, 0, 3.1416}};
    memcpy(r->DH, DH, sizeof(DH));
    /* robotproperty2.m:36-52 ; cap{i}.p is 3x2, columns = endpoints */
    const double cap[6][2][3] = {{{0, 0, 0}, {0, 0, 0}},
                                 {{-0.4, 0, 0}, {0, 0, 0}},
                                 {{-0.03, 0, 0.05}, {-0.03, 0, 0.05}},
                                 {{0, 0, 0}, {0, 0.4, 0}},
                                 {{0, 0, -0.26}, {0, 0, 0.01}},
                                 {{0.05, 0, 0.1107}, {0.18, 0, 0.1107}}};
    memcpy(r->cap, cap, sizeof(cap));
    /* robotproperty2.m:53-54  offset=[3150,8500,330]./1000 */
    r->base[0] = 3150. / 1000;
    r->base[1] = 8500. / 1000;
    r->base[2] = 330. / 1000;
    r->nj = 5;
  } else if (kind == ORC_M16IB) {
    /* robotproperty2.m:68-73 */
    const double DH[6][4] = {{0.5, 0.65, 0.15, 1.5708},   {1.5708, 0, 0.77, 0},       {0, 0, 0.1, 1.5708},
                             {0, 0.74, 0, -1.5708},       {-PI_ / 2, 0, 0, 1.5708},   {PI_, 0.1, 0, 0}};
    memcpy(r->DH, DH, sizeof(DH));
    /* robotproperty2.m:77-94 */
    const double cap[6][2][3] = {{{0, 0, -0.1}, {0, 0, 0.1}},
                                 {{-0.75, 0, -0.15}, {0, 0, -0.15}},
                                 {{-0.03, 0, 0.05}, {-0.03, 0, 0.05}},
                                 {{0, 0, 0}, {0, 0.55, 0}},
                                 {{0, 0, -0.05}, {0, 0, 0.110}},
                                 {{-0.11, 0, 0.09}, {-0.11, 0, 0.09}}};
    memcpy(r->cap, cap, sizeof(cap));
    /* robotproperty2.m:98-99  offset=[3250,8500,0]./1000 */
    r->base[0] = 3250. / 1000;
    r->base[1] = 8500. / 1000;
    r->base[2] = 0. / 1000;
    r->nj = 5;
  } else {
    /* robotproperty2.m:112-128 */
    const double DH[3][4] = {{0, 0, 0.3, 0}, {0, 0, 0.2, 0}, {0, 0, 0, 0}};
    memcpy(r->DH, DH, sizeof(DH));
    r->cap[0][0][0] = 0;   r->cap[0][1][0] = 0.3;
    r->cap[1][0][0] = 0;   r->cap[1][1][0] = 0.2;
    /* robot.T = [0 0 0.3; 0 0 0; 0 0 0] -> T(:,3) = [0.3;0;0] */
    r->T2L[2][0] = 0.3;
    r->nj = 2;
  }
}

/* ------------------------------------------------------------------------------------------
 * CapPos.  Lib/functions/CapPos.m:8-23 :  M{i+1}=M{i}*[R T;0 0 0 1],
 *   pos{i}.p(:,k)=M{i+1}(1:3,1:3)*RoCap{i}.p(:,k)+M{i+1}(1:3,4)+base
 * M is kept as 3x4 (the bottom row is [0 0 0 1] throughout).
 * ---------------------------------------------------------------------------------------- */
static void chain_step(const double M[3][4], const double R[3][3], const double T[3], double Mn[3][4]) {
  for (int a = 0; a < 3; ++a) {
    for (int c = 0; c < 3; ++c) Mn[a][c] = (M[a][0] * R[0][c] + M[a][1] * R[1][c]) + M[a][2] * R[2][c];
    Mn[a][3] = ((M[a][0] * T[0] + M[a][1] * T[1]) + M[a][2] * T[2]) + M[a][3];
  }
}

static void cap_pos_dh(const double DH[][4], int nlink, const double base[3], const double cap[][2][3],
                       double *pos) {
  double M[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
  for (int i = 0; i < nlink; ++i) {
    const double th = DH[i][0], d = DH[i][1], a = DH[i][2], al = DH[i][3];
    const double ct = cos(th), st = sin(th), ca = cos(al), sa = sin(al);
    /* CapPos.m:13-16 */
    const double R[3][3] = {{ct, -st * ca, st * sa}, {st, ct * ca, -ct * sa}, {0, sa, ca}};
    const double T[3] = {a * ct, a * st, d};
    double Mn[3][4];
    chain_step(M, R, T, Mn);
    memcpy(M, Mn, sizeof(Mn));
    for (int k = 0; k < 2; ++k) /* CapPos.m:18-20 */
      for (int a3 = 0; a3 < 3; ++a3)
        pos[(i * 2 + k) * 3 + a3] =
            (((M[a3][0] * cap[i][k][0] + M[a3][1] * cap[i][k][1]) + M[a3][2] * cap[i][k][2]) + M[a3][3]) + base[a3];
  }
}

/* CapPos2.  Lib/2L/CapPos2.m:16-29 : planar rotations, translation robot.T(:,i) */
static void cap_pos_2l(const orc_robot *r, const double *theta, double *pos) {
  double M[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
  for (int i = 1; i <= r->nj; ++i) {
    const double ct = cos(theta[i - 1]), st = sin(theta[i - 1]);
    const double R[3][3] = {{ct, -st, 0}, {st, ct, 0}, {0, 0, 1}};
    const double T[3] = {r->T2L[i][0], r->T2L[i][1], r->T2L[i][2]}; /* T(:,i) with i=2..nlink+1 (1-based) */
    double Mn[3][4];
    chain_step(M, R, T, Mn);
    memcpy(M, Mn, sizeof(Mn));
    for (int k = 0; k < 2; ++k)
      for (int a3 = 0; a3 < 3; ++a3)
        pos[((i - 1) * 2 + k) * 3 + a3] =
            (((M[a3][0] * r->cap[i - 1][k][0] + M[a3][1] * r->cap[i - 1][k][1]) + M[a3][2] * r->cap[i - 1][k][2]) +
             M[a3][3]) + r->base[a3];
  }
}

/* FK as used by the dist_arm_* functions: DH(i,1)=theta(i) (+ M200i joint-2 offset).
 * dist_arm_3D_Heu_2.m:4-18 ; dist_arm_3D_200i_2.m:4-19 (offset :11) ; dist_arm_2L.m:11 */
void orc_cap_pos(const orc_robot *r, const double *theta, double *pos) {
  if (r->kind == ORC_2L) {
    cap_pos_2l(r, theta, pos);
    return;
  }
  double DH[ORC_MAXL][4];
  memcpy(DH, r->DH, sizeof(DH));
  for (int i = 0; i < r->nj; ++i) DH[i][0] = theta[i];
  if (r->kind == ORC_M200I) DH[1][0] = DH[1][0] - PI_ / 2;
  cap_pos_dh(DH, r->nj, r->base, r->cap, pos);
}

/* ------------------------------------------------------------------------------------------
 * distLinSeg.  Lib/functions/distLinSeg.m:23-101 (Lumelsky 1985)
 * ---------------------------------------------------------------------------------------- */
static double fixbound(double v) { /* distLinSeg.m:93-101 */
  if (v < 0) return 0;
  if (v > 1) return 1;
  return v;
}

double orc_dist_lin_seg(const double *p1s, const double *p1e, const double *p2s, const double *p2e, int dim,
                        double *points) {
  double d1[3] = {0, 0, 0}, d2[3] = {0, 0, 0}, d12[3] = {0, 0, 0};
  double D1 = 0, D2 = 0, S1 = 0, S2 = 0, R = 0;
  for (int k = 0; k < dim; ++k) { /* :25-27 */
    d1[k] = p1e[k] - p1s[k];
    d2[k] = p2e[k] - p2s[k];
    d12[k] = p2s[k] - p1s[k];
  }
  for (int k = 0; k < dim; ++k) { /* :29-34 */
    D1 += d1[k] * d1[k];
    D2 += d2[k] * d2[k];
    S1 += d1[k] * d12[k];
    S2 += d2[k] * d12[k];
    R += d1[k] * d2[k];
  }
  const double den = D1 * D2 - R * R; /* :36 */
  double t, u;
  if (D1 == 0 || D2 == 0) { /* :38-54 */
    if (D1 != 0) {
      u = 0;
      t = fixbound(S1 / D1);
    } else if (D2 != 0) {
      t = 0;
      u = fixbound(-S2 / D2);
    } else {
      t = 0;
      u = 0;
    }
  } else if (den == 0) { /* :55-66 */
    t = 0;
    u = -S2 / D2;
    const double uf = fixbound(u);
    if (uf != u) {
      t = fixbound((uf * R + S1) / D1);
      u = uf;
    }
  } else { /* :67-82 */
    t = fixbound((S1 * D2 - S2 * R) / den);
    u = (t * R - S2) / D2;
    const double uf = fixbound(u);
    if (uf != u) {
      t = fixbound((uf * R + S1) / D1);
      u = uf;
    }
  }
  double ss = 0; /* :85  norm(d1*t-d2*u-d12) */
  for (int k = 0; k < dim; ++k) {
    const double v = (d1[k] * t - d2[k] * u) - d12[k];
    ss += v * v;
  }
  if (points) { /* :88  [point1s+d1*t ; point2s+d2*u] */
    for (int k = 0; k < dim; ++k) {
      points[k] = p1s[k] + d1[k] * t;
      points[dim + k] = p2s[k] + d2[k] * u;
    }
  }
  return sqrt(ss);
}

/* link distance with the "negative when axes touch" heuristic.
 * dist_arm_3D_200i_2.m:21-24 / dist_link_Heu.m:18-21 form (points(1:3)).  On M16iB the class path
 * (dist_arm_3D_Heu_2.m:23) subtracts a 3x1 from a 6x1 and MATLAB throws: flagged via *touched. */
/* ---- N3: box obstacles -- an EXTENSION, the reference only has capsule-axis obstacles (obs{j}.l as a line
 * segment; its mesh path Lib/functions/dist_arm_surface.m:44 calls point2surface_dis, which is defined nowhere).  An obstacle
 * record is ORC_OBS_STRIDE = 7 doubles: [l(:,1); l(:,2); kind].  kind 0: capsule axis from l(:,1) to l(:,2).  kind 1
 * (obs{j}.shape = 'box'): the solid axis-aligned box with min corner l(:,1) and max corner l(:,2), e.g. the bounding box of an
 * STL part of map/ (MapFromSTL.m:1-12 reads those in mm; the robot lives in m).
 *
 * orc_dist_seg_box: distance between the segment [ps, pe] and the solid box [lo, hi], and the closest point of the segment.
 * With x(t) = ps + t (pe - ps) and the signed per-axis excess e_k(x) = x_k - hi_k (x_k > hi_k), x_k - lo_k (x_k < lo_k), 0 inside,
 * f(t) = sum_k e_k(x(t))^2 is convex and C1, so g(t) = f'(t)/2 = sum_k e_k d_k is continuous, piecewise linear and
 * nondecreasing with breakpoints where x_k(t) crosses lo_k / hi_k.  The FIRST minimiser over [0,1]: t = 0 if g(0) >= 0, t = 1 if
 * g(1) < 0, else the first root of g, found exactly by evaluating g at the (<= 6) breakpoints inside (0,1), keeping the nearest one
 * on either side of the root and interpolating linearly between them.  A segment that meets the box gives distance 0 at the
 * first such t. */
static double box_excess_dot(const double *ps, const double *d, const double *lo, const double *hi, double t, double *f) {
  double g = 0.0, ss = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double x = ps[k] + d[k] * t;
    const double e = x > hi[k] ? x - hi[k] : (x < lo[k] ? x - lo[k] : 0.0);
    g += e * d[k];
    ss += e * e;
  }
  if (f) *f = ss;
  return g;
}

double orc_dist_seg_box(const double *ps, const double *pe, const double *lo, const double *hi, double *point /*3 or NULL*/) {
  double d[3];
  for (int k = 0; k < 3; ++k) d[k] = pe[k] - ps[k];
  double t;
  const double g0 = box_excess_dot(ps, d, lo, hi, 0.0, NULL);
  if (g0 >= 0.0) {
    t = 0.0;
  } else {
    const double g1 = box_excess_dot(ps, d, lo, hi, 1.0, NULL);
    if (g1 < 0.0) {
      t = 1.0;
    } else {
      double tl = 0.0, gl = g0, th = 1.0, gh = g1;
      for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.0) continue;
        for (int side = 0; side < 2; ++side) {
          const double c = ((side ? hi[k] : lo[k]) - ps[k]) / d[k];
          if (!(c > 0.0 && c < 1.0)) continue;
          const double gc = box_excess_dot(ps, d, lo, hi, c, NULL);
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
  double ss;
  (void)box_excess_dot(ps, d, lo, hi, t, &ss);
  if (point)
    for (int k = 0; k < 3; ++k) point[k] = ps[k] + d[k] * t;
  return sqrt(ss);
}

static double link_dist(const double *pos_i /* 2*3 */, const double *obs6, int *touched) {
  double points[6];
  double dis = obs6[6] == 1.0 ? orc_dist_seg_box(pos_i, pos_i + 3, obs6, obs6 + 3, points)
                              : orc_dist_lin_seg(pos_i, pos_i + 3, obs6, obs6 + 3, 3, points);
  if (fabs(dis) < 0.0001) {
    double ss = 0;
    for (int k = 0; k < 3; ++k) {
      const double v = points[k] - pos_i[3 + k];
      ss += v * v;
    }
    dis = -sqrt(ss);
    if (touched) *touched = 1;
  }
  return dis;
}

/* dist_arm_3D_Heu_2.m:1-30 ; dist_arm_3D_200i_2.m:1-30 ; dist_arm_2L.m:1-23.  linkid is 1-based. */
double orc_dist_arm(const orc_robot *r, const double *theta, const double *obs6, int *linkid, int *touched) {
  double pos[ORC_MAXL * 6];
  orc_cap_pos(r, theta, pos);
  double d = INFINITY;
  int id = 0;
  for (int i = 0; i < r->nj; ++i) {
    const double dis = link_dist(pos + 6 * i, obs6, touched);
    if (dis < d) { /* strict <: first minimal link wins (:25-28) */
      d = dis;
      id = i + 1;
    }
  }
  if (linkid) *linkid = id;
  return d;
}

/* dist_link_Heu.m:1-26 ; dist_link_200i.m:1-25 : distance of the single link `linkid` (1-based) */
double orc_dist_link(const orc_robot *r, const double *theta, const double *obs6, int linkid, int *touched) {
  double pos[ORC_MAXL * 6];
  orc_cap_pos(r, theta, pos);
  return link_dist(pos + 6 * (linkid - 1), obs6, touched);
}

/* num_jac.  Lib/functions/num_jac.m:1-17, eps = 1e-5.
 * NOTE (faithful quirk): xp(i) is left at x(i)-eps/2 after its column is done (:13-14, never reset),
 * so column i is evaluated with all earlier joints shifted by -eps/2. */
void orc_num_jac(const orc_robot *r, const double *theta, const double *obs6, double *grad, int *touched) {
  const double eps = 1e-5;
  double xp[ORC_MAXL];
  (void)orc_dist_arm(r, theta, obs6, NULL, touched); /* y = f(x), :2 */
  for (int i = 0; i < r->nj; ++i) xp[i] = theta[i];
  for (int i = 0; i < r->nj; ++i) {
    xp[i] = theta[i] + eps / 2;
    const double yhi = orc_dist_arm(r, xp, obs6, NULL, touched);
    xp[i] = theta[i] - eps / 2;
    const double ylo = orc_dist_arm(r, xp, obs6, NULL, touched);
    grad[i] = (yhi - ylo) / eps;
  }
}

/* ------------------------------------------------------------------------------------------
 * DERIVEST, defaults: DerivativeOrder 1, MethodOrder 4, central, RombergTerms 2,
 * StepRatio 2.0000001, MaxStep 100.  DERIVESTsuite/DERIVESTsuite/derivest.m:193-468,
 * rombextrap :475-530, vec2mat :536-545, fdamat :551-580.
 * ---------------------------------------------------------------------------------------- */
#define DV_NDEL 26
#define DV_NE 23
#define DV_NEST 19

static void derivest_constants(double sr, double fdarule[2], double rmat[4][3], double pinv1[4] /* first row of pinv */,
                               double qmat[4][3], double rr[3][3], double *cov11) {
  /* fdamat(sr,1,2) :568-572 -> mat(i,j)=c(j)*srinv^((i-1)*(2j-1)), c=[1,1/6];  fdarule=[1 0]/mat (:282) */
  const double srinv = 1.0 / sr;
  const double m00 = 1.0, m01 = 1.0 / 6.0, m10 = srinv, m11 = (1.0 / 6.0) * pow(srinv, 3);
  /* solve x*mat=[1 0]  <=> mat' x' = [1;0] by Gaussian elimination with partial pivoting */
  {
    double a[2][3] = {{m00, m10, 1.0}, {m01, m11, 0.0}};
    if (fabs(a[1][0]) > fabs(a[0][0])) {
      for (int k = 0; k < 3; ++k) { double t = a[0][k]; a[0][k] = a[1][k]; a[1][k] = t; }
    }
    const double f = a[1][0] / a[0][0];
    a[1][1] -= f * a[0][1];
    a[1][2] -= f * a[0][2];
    const double x2 = a[1][2] / a[1][1];
    const double x1 = (a[0][2] - a[0][1] * x2) / a[0][0];
    fdarule[0] = x1;
    fdarule[1] = x2;
  }
  /* rombextrap :493-499 with rombexpon=[4 6] (:431) */
  const double ex[2] = {4, 6};
  for (int i = 0; i < 4; ++i) {
    rmat[i][0] = 1.0;
    for (int j = 0; j < 2; ++j) rmat[i][1 + j] = (i == 0) ? 1.0 : pow(srinv, i * ex[j]);
  }
  /* economy QR (:510) by modified Gram-Schmidt with re-orthogonalisation (4x3, well conditioned) */
  double q[4][3];
  for (int j = 0; j < 3; ++j) {
    double v[4];
    for (int i = 0; i < 4; ++i) v[i] = rmat[i][j];
    for (int k = 0; k < 3; ++k) rr[k][j] = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int k = 0; k < j; ++k) {
        double dot = 0;
        for (int i = 0; i < 4; ++i) dot += q[i][k] * v[i];
        rr[k][j] += dot;
        for (int i = 0; i < 4; ++i) v[i] -= dot * q[i][k];
      }
    double nn = 0;
    for (int i = 0; i < 4; ++i) nn += v[i] * v[i];
    nn = sqrt(nn);
    rr[j][j] = nn;
    for (int i = 0; i < 4; ++i) q[i][j] = v[i] / nn;
  }
  memcpy(qmat, q, sizeof(q));
  /* rinv = rromb\eye(3) ; cov1 = sum(rinv.^2,2) (:523-524) */
  double rinv[3][3] = {{0}};
  for (int c = 0; c < 3; ++c) {
    for (int i = 2; i >= 0; --i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = i + 1; k < 3; ++k) s -= rr[i][k] * rinv[k][c];
      rinv[i][c] = s / rr[i][i];
    }
  }
  *cov11 = rinv[0][0] * rinv[0][0] + rinv[0][1] * rinv[0][1] + rinv[0][2] * rinv[0][2];
  /* first row of rromb\qromb' */
  for (int i = 0; i < 4; ++i) pinv1[i] = rinv[0][0] * q[i][0] + rinv[0][1] * q[i][1] + rinv[0][2] * q[i][2];
}

void orc_derivest(orc_fun1 fun, void *ctx, double x0, double *der, double *errest, double *finaldelta) {
  const double sr = 2.0000001;            /* :203 */
  const double h = (x0 > 0.02) ? x0 : 0.02; /* par.NominalStep = max(x0,0.02) :229 */
  double delta[DV_NDEL];
  for (int k = 0; k < DV_NDEL; ++k) delta[k] = 100.0 * pow(sr, (double)(-k)); /* :238 */

  double fdarule[2], rmat[4][3], pinv1[4], q[4][3], rr[3][3], cov11;
  derivest_constants(sr, fdarule, rmat, pinv1, q, rr, &cov11);

  double f_del[DV_NDEL];
  for (int j = 0; j < DV_NDEL; ++j) { /* :366-371, :376 */
    const double fp = fun(x0 + h * delta[j], ctx);
    const double fm = fun(x0 - h * delta[j], ctx);
    f_del[j] = (fp - fm) / 2;
  }
  /* der_init = vec2mat(f_del,ne,nfda)*fdarule.' (:415) ./ (h*delta(1:ne)) (:418) */
  double der_init[DV_NE];
  for (int i = 0; i < DV_NE; ++i)
    der_init[i] = (f_del[i] * fdarule[0] + f_del[i + 1] * fdarule[1]) / (h * delta[i]);

  /* rombextrap :512-526 : rhs(i,j)=der_init(i+j), 4 x 19 */
  double der_romb[DV_NEST], errors[DV_NEST];
  for (int j = 0; j < DV_NEST; ++j) {
    double qtr[3], coef[3];
    for (int c = 0; c < 3; ++c) {
      qtr[c] = 0;
      for (int i = 0; i < 4; ++i) qtr[c] += q[i][c] * der_init[i + j];
    }
    for (int i = 2; i >= 0; --i) {
      double s = qtr[i];
      for (int k = i + 1; k < 3; ++k) s -= rr[i][k] * coef[k];
      coef[i] = s / rr[i][i];
    }
    der_romb[j] = coef[0];
    double ss = 0;
    for (int i = 0; i < 4; ++i) {
      const double res = der_init[i + j] - ((rmat[i][0] * coef[0] + rmat[i][1] * coef[1]) + rmat[i][2] * coef[2]);
      ss += res * res;
    }
    errors[j] = sqrt(ss) * 12.7062047361747 * sqrt(cov11); /* :525 */
  }
  (void)pinv1;
  /* :442-462 : stable ascending sort, delete ranks [1 2 nest-1 nest], min error */
  int tags[DV_NEST];
  for (int j = 0; j < DV_NEST; ++j) tags[j] = j;
  for (int a = 1; a < DV_NEST; ++a) { /* stable insertion sort; NaN sorts last like MATLAB */
    const int t = tags[a];
    int b = a - 1;
    while (b >= 0 && (der_romb[tags[b]] > der_romb[t] || (isnan(der_romb[tags[b]]) && !isnan(der_romb[t])))) {
      tags[b + 1] = tags[b];
      --b;
    }
    tags[b + 1] = t;
  }
  int best = -1;
  double best_err = 0;
  for (int a = 2; a < DV_NEST - 2; ++a) {
    const double e = errors[tags[a]];
    if (best < 0 || e < best_err) {
      best = a;
      best_err = e;
    }
  }
  if (der) *der = der_romb[tags[best]];
  if (errest) *errest = best_err;
  if (finaldelta) *finaldelta = h * delta[tags[best]];
}

typedef struct {
  const orc_robot *r;
  const double *obs6;
  double theta[ORC_MAXL];
  int s, linkid;
  int *touched;
} dl_ctx;

static double dl_fun(double x, void *vctx) {
  dl_ctx *c = (dl_ctx *)vctx;
  double th[ORC_MAXL];
  for (int i = 0; i < c->r->nj; ++i) th[i] = c->theta[i];
  th[c->s] = x;
  return orc_dist_link(c->r, th, c->obs6, c->linkid, c->touched);
}

/* M16iB/main_CFS.m:234-237 : Diff(s)=derivest(@(x) dist_link_Heu([theta(1:s-1);x;theta(s+1:end)],...,linkid),theta(s),'Vectorized','no') */
void orc_derivest_grad(const orc_robot *r, const double *theta, const double *obs6, int linkid, double *grad,
                       int *touched) {
  dl_ctx c;
  c.r = r;
  c.obs6 = obs6;
  c.linkid = linkid;
  c.touched = touched;
  for (int i = 0; i < r->nj; ++i) c.theta[i] = theta[i];
  for (int s = 0; s < r->nj; ++s) {
    c.s = s;
    orc_derivest(dl_fun, &c, theta[s], &grad[s], NULL, NULL);
  }
}

static double nf_exp(double x, void *c) { (void)c; return exp(x); }
static double nf_sin(double x, void *c) { (void)c; return sin(x); }
static double nf_sinh(double x, void *c) { (void)c; return sinh(x); }
static double nf_log(double x, void *c) { (void)c; return log(x); }
static double nf_cube(double x, void *c) { (void)c; return x * x * x + x * x * x * x; }
void orc_derivest_named(int which, double x0, double *der, double *errest, double *finaldelta) {
  orc_fun1 f = which == 0 ? nf_exp : which == 1 ? nf_sin : which == 2 ? nf_sinh : which == 3 ? nf_log : nf_cube;
  orc_derivest(f, NULL, x0, der, errest, finaldelta);
}

/* ------------------------------------------------------------------------------------------
 * Cost builder.  main_FANUC.m:64-103 (also RRTstar_CFS.m:124-163, main_2L.m:69-95).
 * State per step = [theta(nj); omega(nj)];  A=[I dt I;0 I], B=[dt^2/2 I; dt I] (robotproperty2.m:136-139).
 * ---------------------------------------------------------------------------------------- */
void orc_build_cost(int nj, int H, double dt, const double *Q, const double *Rblk, double r_scale, double stage_w,
                    double term_w, double *Aaug, double *Baug, double *unused, double *QQ) {
  (void)unused;
  const int ns = 2 * nj, n = nj * H, N = ns * H;
  /* Aaug=[A^1;...;A^H] (:79) ; A^i = [I i*dt*I; 0 I] */
  if (Aaug) {
    memset(Aaug, 0, sizeof(double) * N * ns);
    for (int i = 1; i <= H; ++i)
      for (int k = 0; k < nj; ++k) {
        Aaug[((i - 1) * ns + k) + (size_t)N * k] = 1.0;
        Aaug[((i - 1) * ns + k) + (size_t)N * (nj + k)] = i * dt;
        Aaug[((i - 1) * ns + nj + k) + (size_t)N * (nj + k)] = 1.0;
      }
  }
  /* Baug block (i,j)=A^(i-j)*B, j<=i (:84-86) */
  double *Bg = Baug ? Baug : (double *)calloc((size_t)N * n, sizeof(double));
  memset(Bg, 0, sizeof(double) * (size_t)N * n);
  for (int i = 1; i <= H; ++i)
    for (int j = 1; j <= i; ++j)
      for (int k = 0; k < nj; ++k) {
        Bg[((i - 1) * ns + k) + (size_t)N * ((j - 1) * nj + k)] = 0.5 * dt * dt + ((i - j) * dt) * dt;
        Bg[((i - 1) * ns + nj + k) + (size_t)N * ((j - 1) * nj + k)] = dt;
      }
  /* QB = Qaug*Baug, Qaug = blkdiag(Q*0.1,...,Q*10000) (:80-83) */
  double *QB = (double *)calloc((size_t)N * n, sizeof(double));
  for (int i = 1; i <= H; ++i) {
    const double w = (i == H) ? term_w : stage_w;
    for (int c = 0; c < n; ++c)
      for (int a = 0; a < ns; ++a) {
        double s = 0;
        for (int b = 0; b < ns; ++b) s += (Q[a + ns * b] * w) * Bg[((i - 1) * ns + b) + (size_t)N * c];
        QB[((i - 1) * ns + a) + (size_t)N * c] = s;
      }
  }
  /* QQ = Baug'*Qaug*Baug + R.*r_scale, R = blkdiag(Rblk) with the identity elsewhere, R=R+R' (:88-97) */
  for (int a = 0; a < n; ++a)
    for (int c = 0; c <= a; ++c) {
      double s = 0;
      for (int k = 0; k < N; ++k) s += Bg[k + (size_t)N * a] * QB[k + (size_t)N * c];
      QQ[a + (size_t)n * c] = s;
      QQ[c + (size_t)n * a] = s;
    }
  for (int i = 0; i < H; ++i)
    for (int a = 0; a < nj; ++a)
      for (int c = 0; c < nj; ++c)
        QQ[(i * nj + a) + (size_t)n * (i * nj + c)] += (Rblk[a + nj * c] + Rblk[c + nj * a]) * r_scale;
  free(QB);
  if (!Baug) free(Bg);
}

/* ff = ((Aaug*x0-gaug)'*Qaug*Baug)' ; caug = (Aaug*x0-gaug)'*Qaug*(Aaug*x0-gaug)  (main_FANUC.m:98-103) */
void orc_build_ff(int nj, int H, const double *Q, double stage_w, double term_w, const double *Aaug,
                  const double *Baug, const double *x0, const double *gaug, double *ff, double *caug) {
  const int ns = 2 * nj, n = nj * H, N = ns * H;
  double *e = (double *)malloc(sizeof(double) * N), *qe = (double *)malloc(sizeof(double) * N);
  for (int k = 0; k < N; ++k) {
    double s = 0;
    for (int c = 0; c < ns; ++c) s += Aaug[k + (size_t)N * c] * x0[c];
    e[k] = s - gaug[k];
  }
  double cc = 0;
  for (int i = 0; i < H; ++i) {
    const double w = (i == H - 1) ? term_w : stage_w;
    for (int a = 0; a < ns; ++a) {
      double s = 0;
      for (int b = 0; b < ns; ++b) s += e[i * ns + b] * (Q[b + ns * a] * w);
      qe[i * ns + a] = s;
    }
  }
  for (int k = 0; k < N; ++k) cc += qe[k] * e[k];
  for (int c = 0; c < n; ++c) {
    double s = 0;
    for (int k = 0; k < N; ++k) s += qe[k] * Baug[k + (size_t)N * c];
    ff[c] = s;
  }
  if (caug) *caug = cc;
  free(e);
  free(qe);
}

/* ------------------------------------------------------------------------------------------
 * get_con.  Lib/CFS_FANUC.m:101-135 (twin Lib/PSGCFS_FANUC.m:145-184; script twin
 * M16iB/main_CFS.m:225-257 = DERIVEST gradients, no velocity rows).
 * Row order per (obstacle j, step i): 1 obstacle row, nj rows +Baug_w, nj rows -Baug_w.
 * Baug blocks are closed-form: theta rows (i,j) = (0.5+(i-j))dt^2 I, omega rows = dt I, j<=i.
 * ---------------------------------------------------------------------------------------- */
int orc_get_con(const orc_robot *r, const orc_cfg *c, const double *x0, const double *xcur, const double *u,
                double *Ainq, double *binq, double *dist, int *linkid, double *grad, int *touched) {
  const int nj = r->nj, ns = 2 * nj, H = c->H, n = nj * H;
  const double dt = r->dt;
  const int rows_per = c->lim ? 1 + 2 * nj : 1;
  int row = 0;
  for (int j = 0; j < c->nobs; ++j) {
    const double *obs6 = c->obs + ORC_OBS_STRIDE * j;
    for (int i = 1; i <= H; ++i) {
      const double *theta = xcur + ns * (i - 1); /* :114 */
      int lid = 0;
      const double distance = orc_dist_arm(r, theta, obs6, &lid, touched); /* :115 */
      const double I = distance - c->margin[j];                             /* :117 */
      double g[ORC_MAXL];
      if (c->grad == 0)
        orc_num_jac(r, theta, obs6, g, touched); /* :118 */
      else
        orc_derivest_grad(r, theta, obs6, lid, g, touched); /* M16iB/main_CFS.m:234-237 */
      if (dist) dist[j * H + (i - 1)] = distance;
      if (linkid) linkid[j * H + (i - 1)] = lid;
      if (grad)
        for (int k = 0; k < nj; ++k) grad[(j * H + (i - 1)) * nj + k] = g[k];
      /* l=-Diff'*Bj(1:njoint,:) ; s=I-Diff'*Bj(1:njoint,:)*u  (:119-121) */
      double *l = Ainq + (size_t)row * n;
      double lu = 0;
      for (int jj = 1; jj <= H; ++jj)
        for (int k = 0; k < nj; ++k) {
          const double b = (jj <= i) ? (0.5 * dt * dt + ((i - jj) * dt) * dt) : 0.0;
          const double v = g[k] * b; /* (Diff'*Bj)(col) : single non-zero term per column */
          l[(jj - 1) * nj + k] = -v;
          lu += v * u[(jj - 1) * nj + k];
        }
      binq[row] = I - lu;
      ++row;
      if (c->lim) { /* :126-129 */
        for (int sgn = 0; sgn < 2; ++sgn)
          for (int k = 0; k < nj; ++k) {
            double *a = Ainq + (size_t)row * n;
            for (int jj = 1; jj <= H; ++jj)
              for (int kk = 0; kk < nj; ++kk) a[(jj - 1) * nj + kk] = (kk == k && jj <= i) ? (sgn ? -dt : dt) : 0.0;
            /* Aaug_w(i,:)*xR(:,1) = omega0(k) */
            const double aw = x0[nj + k];
            binq[row] = sgn ? c->lim[k] + aw : c->lim[k] - aw;
            ++row;
          }
      }
    }
  }
  (void)rows_per;
  return row;
}

/* ------------------------------------------------------------------------------------------
 * Strictly convex QP: Goldfarb & Idnani (1983) dual active-set method, dense.
 * Stands in for quadprog (Lib/CFS_FANUC.m:85, Lib/PSGCFS_FANUC.m:120); the optimum is unique.
 * ---------------------------------------------------------------------------------------- */
int orc_chol_J0(int n, const double *G, double *J0) {
  /* G = L L' (lower), J0 = L^{-T}; column-major */
  double *L = (double *)malloc(sizeof(double) * n * n);
  memcpy(L, G, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double s = L[j + (size_t)n * j];
    for (int k = 0; k < j; ++k) s -= L[j + (size_t)n * k] * L[j + (size_t)n * k];
    if (!(s > 0)) {
      free(L);
      return 3;
    }
    const double d = sqrt(s);
    L[j + (size_t)n * j] = d;
    for (int i = j + 1; i < n; ++i) {
      double t = L[i + (size_t)n * j];
      for (int k = 0; k < j; ++k) t -= L[i + (size_t)n * k] * L[j + (size_t)n * k];
      L[i + (size_t)n * j] = t / d;
    }
  }
  /* Linv (lower) by forward substitution per column, then J0 = Linv' */
  double *Li = (double *)calloc((size_t)n * n, sizeof(double));
  for (int c = 0; c < n; ++c) {
    for (int i = c; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= L[i + (size_t)n * k] * Li[k + (size_t)n * c];
      Li[i + (size_t)n * c] = s / L[i + (size_t)n * i];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < n; ++c) J0[c + (size_t)n * i] = Li[i + (size_t)n * c];
  free(L);
  free(Li);
  return 0;
}

static void givens(double a, double b, double *cc, double *ss) {
  if (b == 0) {
    *cc = 1;
    *ss = 0;
  } else {
    const double h = hypot(a, b);
    *cc = a / h;
    *ss = b / h;
  }
}

int orc_qp_gi(int n, int m, const double *J0, const double *xunc, const double *C, const double *d, double *x,
              double *lam_out, int *iters_out) {
  return orc_qp_gi2(n, m, J0, xunc, C, d, x, lam_out, iters_out, NULL);
}

int orc_qp_gi2(int n, int m, const double *J0, const double *xunc, const double *C, const double *d, double *x,
               double *lam_out, int *iters_out, int *qmax_out) {
  int ret = 0, qmax = 0;
  double *J = (double *)malloc(sizeof(double) * n * n);
  double *R = (double *)calloc((size_t)n * n, sizeof(double)); /* upper-triangular, col-major, ld n */
  double *dv = (double *)malloc(sizeof(double) * n), *z = (double *)malloc(sizeof(double) * n);
  double *rv = (double *)malloc(sizeof(double) * n), *uu = (double *)calloc(n + 1, sizeof(double));
  double *cnorm = (double *)malloc(sizeof(double) * (m > 0 ? m : 1));
  int *A = (int *)malloc(sizeof(int) * (n + 1));
  char *inA = (char *)calloc(m > 0 ? m : 1, 1);
  int q = 0, iters = 0;
  memcpy(J, J0, sizeof(double) * n * n);
  memcpy(x, xunc, sizeof(double) * n);
  for (int i = 0; i < m; ++i) {
    double s = 0;
    for (int k = 0; k < n; ++k) s += C[(size_t)i * n + k] * C[(size_t)i * n + k];
    cnorm[i] = sqrt(s);
  }
  const int max_iters = 20 * (m + n) + 100;
  for (;;) {
    /* step 1: most violated constraint (normalised) */
    int p = -1;
    double worst = 0;
    for (int i = 0; i < m; ++i) {
      if (inA[i] || cnorm[i] == 0) continue;
      double s = d[i];
      for (int k = 0; k < n; ++k) s -= C[(size_t)i * n + k] * x[k];
      const double tol = 1e-12 * (1.0 + fabs(d[i]) / cnorm[i]);
      const double v = s / cnorm[i];
      if (v < -tol && (p < 0 || v < worst)) {
        p = i;
        worst = v;
      }
    }
    if (p < 0) break; /* optimal */
    const double *cp = C + (size_t)p * n;
    uu[q] = 0;
    for (;;) { /* step 2 */
      if (++iters > max_iters) {
        ret = ORC_NUMERICAL;
        goto done;
      }
      /* dv = J' n+ with n+ = -c_p */
      double dn2 = 0, d22 = 0;
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < n; ++k) s -= J[k + (size_t)n * j] * cp[k];
        dv[j] = s;
        dn2 += s * s;
        if (j >= q) d22 += s * s;
      }
      const int dependent = !(d22 > 1e-20 * dn2);
      /* z = J2 d2 ; r = R^{-1} d1 */
      for (int k = 0; k < n; ++k) z[k] = 0;
      if (!dependent)
        for (int j = q; j < n; ++j)
          for (int k = 0; k < n; ++k) z[k] += J[k + (size_t)n * j] * dv[j];
      for (int i = q - 1; i >= 0; --i) {
        double s = dv[i];
        for (int k = i + 1; k < q; ++k) s -= R[i + (size_t)n * k] * rv[k];
        rv[i] = s / R[i + (size_t)n * i];
      }
      /* step lengths */
      double t1 = INFINITY;
      int l = -1;
      for (int k = 0; k < q; ++k)
        if (rv[k] > 0) {
          const double t = uu[k] / rv[k];
          if (t < t1) {
            t1 = t;
            l = k;
          }
        }
      double sp = d[p];
      for (int k = 0; k < n; ++k) sp -= cp[k] * x[k];
      double t2 = INFINITY;
      double ztn = 0;
      if (!dependent) {
        for (int k = 0; k < n; ++k) ztn -= z[k] * cp[k]; /* z' n+ > 0 */
        t2 = -sp / ztn;
        if (t2 < 0) t2 = 0;
      }
      if (t1 == INFINITY && t2 == INFINITY) {
        ret = ORC_QP_INFEASIBLE;
        goto done;
      }
      const double t = (t1 < t2) ? t1 : t2;
      if (t2 == INFINITY) { /* dual step only */
        for (int k = 0; k < q; ++k) uu[k] -= t * rv[k];
        uu[q] += t;
      } else {
        for (int k = 0; k < n; ++k) x[k] += t * z[k];
        for (int k = 0; k < q; ++k) uu[k] -= t * rv[k];
        uu[q] += t;
      }
      if (t2 <= t1) {
        /* full step: add constraint p. Givens on dv from the bottom, applied to J columns */
        for (int j = n - 1; j > q; --j) {
          double cc, ss;
          givens(dv[j - 1], dv[j], &cc, &ss);
          if (ss == 0) continue;
          dv[j - 1] = cc * dv[j - 1] + ss * dv[j];
          dv[j] = 0;
          for (int k = 0; k < n; ++k) {
            const double a = J[k + (size_t)n * (j - 1)], b = J[k + (size_t)n * j];
            J[k + (size_t)n * (j - 1)] = cc * a + ss * b;
            J[k + (size_t)n * j] = -ss * a + cc * b;
          }
        }
        for (int i = 0; i <= q; ++i) R[i + (size_t)n * q] = dv[i];
        A[q] = p;
        inA[p] = 1;
        ++q;
        if (q > qmax) qmax = q;
        break; /* back to step 1 */
      }
      /* partial step: drop constraint A[l] */
      inA[A[l]] = 0;
      for (int k = l; k < q - 1; ++k) {
        A[k] = A[k + 1];
        uu[k] = uu[k + 1];
        for (int i = 0; i <= k + 1; ++i) R[i + (size_t)n * k] = R[i + (size_t)n * (k + 1)];
      }
      uu[q - 1] = uu[q];
      uu[q] = 0;
      --q;
      for (int j = l; j < q; ++j) { /* restore triangular R; rotate rows j,j+1 and J columns j,j+1 */
        double cc, ss;
        givens(R[j + (size_t)n * j], R[(j + 1) + (size_t)n * j], &cc, &ss);
        if (ss == 0) continue;
        for (int k = j; k < q; ++k) {
          const double a = R[j + (size_t)n * k], b = R[(j + 1) + (size_t)n * k];
          R[j + (size_t)n * k] = cc * a + ss * b;
          R[(j + 1) + (size_t)n * k] = -ss * a + cc * b;
        }
        for (int k = 0; k < n; ++k) {
          const double a = J[k + (size_t)n * j], b = J[k + (size_t)n * (j + 1)];
          J[k + (size_t)n * j] = cc * a + ss * b;
          J[k + (size_t)n * (j + 1)] = -ss * a + cc * b;
        }
      }
      for (int i = 0; i < n; ++i) R[i + (size_t)n * q] = 0;
    }
  }
done:
  if (lam_out) {
    for (int i = 0; i < m; ++i) lam_out[i] = 0;
    for (int k = 0; k < q; ++k) lam_out[A[k]] = uu[k];
  }
  if (iters_out) *iters_out = iters;
  if (qmax_out) *qmax_out = qmax;
  free(J); free(R); free(dv); free(z); free(rv); free(uu); free(cnorm); free(A); free(inA);
  return ret;
}

/* max-norm KKT residual of (x,lam) for min 1/2x'Gx+a'x s.t. Cx<=d: stationarity, primal, dual, complementarity */
double orc_kkt_residual(int n, int m, const double *G, const double *a, const double *C, const double *d,
                        const double *x, const double *lam) {
  double worst = 0;
  for (int i = 0; i < n; ++i) {
    double s = a[i];
    for (int k = 0; k < n; ++k) s += G[i + (size_t)n * k] * x[k];
    for (int k = 0; k < m; ++k) s += C[(size_t)k * n + i] * lam[k];
    if (fabs(s) > worst) worst = fabs(s);
  }
  for (int k = 0; k < m; ++k) {
    double s = d[k];
    for (int i = 0; i < n; ++i) s -= C[(size_t)k * n + i] * x[i];
    if (-s > worst) worst = -s;
    if (-lam[k] > worst) worst = -lam[k];
    if (fabs(s * lam[k]) > worst) worst = fabs(s * lam[k]);
  }
  return worst;
}

/* ------------------------------------------------------------------------------------------
 * CFS_FANUC.optimizer (Lib/CFS_FANUC.m:62-79) / PSGCFS_FANUC.optimizer (Lib/PSGCFS_FANUC.m:65-82)
 * with EVAL (Lib/EVAL.m:51-73).
 * ---------------------------------------------------------------------------------------- */
static double get_cost(int n, const double *QQ, const double *ff, double caug, const double *u) { /* EVAL.m:51-53 */
  double quad = 0, lin = 0;
  for (int c = 0; c < n; ++c) {
    double s = 0;
    for (int k = 0; k < n; ++k) s += u[k] * QQ[k + (size_t)n * c];
    quad += s * u[c];
    lin += ff[c] * u[c];
  }
  return (0.5 * quad + lin) + caug;
}

static void rollout(int nj, int H, double dt, const double *x0, const double *u, double *x) { /* CFS_FANUC.m:90-94 */
  double cur[2 * ORC_MAXL];
  for (int k = 0; k < 2 * nj; ++k) cur[k] = x0[k];
  for (int i = 0; i < H; ++i) {
    for (int k = 0; k < nj; ++k) {
      const double uk = u[i * nj + k];
      const double th = (cur[k] + dt * cur[nj + k]) + (0.5 * dt * dt) * uk;
      const double om = cur[nj + k] + dt * uk;
      cur[k] = th;
      cur[nj + k] = om;
    }
    for (int k = 0; k < 2 * nj; ++k) x[i * 2 * nj + k] = cur[k];
  }
}

int orc_cfs_solve(const orc_robot *r, const orc_cfg *c, const double *J0in, const double *x0, const double *ff,
                  double caug, const double *xref, const double *noise, double *u, double *x, double *cost_hist,
                  double *e_u_hist, int *iters_out, int *qp_iters_out) {
  const int nj = r->nj, ns = 2 * nj, H = c->H, n = nj * H, N = ns * H;
  const int mcon = c->nobs * H * (c->lim ? 1 + 2 * nj : 1);
  const int use_bounds = (c->solver == 0 && c->max_input != NULL);
  const int m = mcon + (use_bounds ? 2 * n : 0);
  int status = ORC_MAX_ITER, touched = 0, qp_total = 0, qmax_all = 0;
  double *J0 = NULL, *J0own = NULL;
  if (c->solver == 0) {
    if (J0in)
      J0 = (double *)J0in;
    else {
      J0own = (double *)malloc(sizeof(double) * n * n);
      if (orc_chol_J0(n, c->QQ, J0own)) {
        free(J0own);
        *iters_out = 0;
        return ORC_NUMERICAL;
      }
      J0 = J0own;
    }
  } else {
    J0own = (double *)calloc((size_t)n * n, sizeof(double)); /* projection Hessian = I (PSGCFS_FANUC.m:117) */
    for (int i = 0; i < n; ++i) J0own[i + (size_t)n * i] = 1.0;
    J0 = J0own;
  }
  double *Cmat = (double *)calloc((size_t)m * n, sizeof(double)), *dvec = (double *)malloc(sizeof(double) * m);
  double *xold = (double *)malloc(sizeof(double) * N), *xunc = (double *)malloc(sizeof(double) * n);
  double *unew = (double *)malloc(sizeof(double) * n), *tmp = (double *)malloc(sizeof(double) * n);
  for (int k = 0; k < n; ++k) u[k] = 0;             /* CFS_FANUC.m:56 */
  memcpy(x, xref, sizeof(double) * N);              /* CFS_FANUC.m:55 */
  for (int k = 0; k < N; ++k) xold[k] = 1.0;        /* EVAL.m:47 */
  if (use_bounds) {                                 /* lb/ub as rows: u<=MAX, -u<=MAX (CFS_FANUC.m:85) */
    for (int k = 0; k < n; ++k) {
      Cmat[(size_t)(mcon + k) * n + k] = 1.0;
      dvec[mcon + k] = c->max_input[k];
      Cmat[(size_t)(mcon + n + k) * n + k] = -1.0;
      dvec[mcon + n + k] = c->max_input[k];
    }
  }
  if (c->solver == 0) { /* unconstrained minimiser -QQ^{-1} ff = -J0 J0' ff, constant over the outer loop */
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += J0[k + (size_t)n * j] * ff[k];
      tmp[j] = s;
    }
    for (int k = 0; k < n; ++k) {
      double s = 0;
      for (int j = 0; j < n; ++j) s += J0[k + (size_t)n * j] * tmp[j];
      xunc[k] = -s;
    }
  }
  double cost_new = get_cost(n, c->QQ, ff, caug, u); /* CFS_FANUC.m:63 */
  double cost_old = 100000;                          /* EVAL.m:29 */
  int iter_O = 1, done_iters = 0;
  for (;;) {
    /* EVAL.stop_outer :61-73 */
    double dd = 0;
    for (int k = 0; k < N; ++k) dd += (x[k] - xold[k]) * (x[k] - xold[k]);
    int stop = 0;
    if (sqrt(dd) < c->eps_outer) {
      stop = 1;
      status = ORC_OK_CONVERGED;
    }
    if (iter_O > c->max_outer) {
      if (!stop) status = ORC_MAX_ITER;
      stop = 1;
    }
    if (stop) break;
    orc_get_con(r, c, x0, x, u, Cmat, dvec, NULL, NULL, NULL, &touched);
    int qit = 0, rc;
    if (c->solver == 0) {
      cost_old = cost_new; /* CFS_FANUC.m:67 */
      int qm = 0;
      rc = orc_qp_gi2(n, m, J0, xunc, Cmat, dvec, unew, NULL, &qit, &qm);
      if (qm > qmax_all) qmax_all = qm;
      if (rc == 0) memcpy(xold, x, sizeof(double) * N); /* CFS_FANUC.m:88 */
    } else {
      /* inner_PSG_5 (PSGCFS_FANUC.m:86-103): MAX_I_ITER=1 -> one step unless |cost_new-cost_old|<epsilon_I */
      int did = 0;
      rc = 0;
      if (!(fabs(cost_new - cost_old) < 1e-4)) {
        cost_old = cost_new;
        /* PSG_update_arm :106-112 */
        for (int k = 0; k < n; ++k) {
          double s = 0;
          for (int j = 0; j < n; ++j) s += c->QQ[k + (size_t)n * j] * u[j];
          const double nz = noise ? noise[(size_t)(iter_O - 1) * n + k] : 0.0;
          xunc[k] = u[k] - c->alpha * ((s + ff[k]) + 10 * nz / ((double)iter_O * iter_O + 1));
        }
        int qm = 0;
        rc = orc_qp_gi2(n, m, J0, xunc, Cmat, dvec, unew, NULL, &qit, &qm); /* Projection :115-128 */
        if (qm > qmax_all) qmax_all = qm;
        did = 1;
      }
      if (!did) memcpy(unew, u, sizeof(double) * n);
      /* NB: PSGCFS never updates eval.x_old (it stays ones) -> runs MAX_O_ITER iterations */
    }
    qp_total += qit;
    if (rc != 0) {
      status = rc;
      break;
    }
    if (e_u_hist) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += (u[k] - unew[k]) * (u[k] - unew[k]);
      e_u_hist[done_iters] = sqrt(s); /* EVAL.m:58 */
    }
    memcpy(u, unew, sizeof(double) * n);
    rollout(nj, H, r->dt, x0, u, x);
    cost_new = get_cost(n, c->QQ, ff, caug, u);
    cost_hist[done_iters] = cost_new; /* EVAL.m:56 */
    ++done_iters;
    ++iter_O;
  }
  *iters_out = done_iters;
  if (qp_iters_out) {
    qp_iters_out[0] = qp_total;
    qp_iters_out[1] = qmax_all;
  }
  free(Cmat); free(dvec); free(xold); free(xunc); free(unew); free(tmp);
  if (J0own) free(J0own);
  return status | (touched ? ORC_FLAG_TOUCH : 0);
}

void orc_cfs_solve_batch(const orc_robot *r, const orc_cfg *c, int B, int nthreads, const double *x0,
                         const double *ff, const double *caug, const double *xref, const double *noise, double *u,
                         double *x, double *cost_hist, double *e_u_hist, int *iters, int *status) {
  orc_cfs_solve_batch2(r, c, B, nthreads, x0, ff, caug, xref, noise, u, x, cost_hist, e_u_hist, iters, status, NULL);
}

void orc_cfs_solve_batch2(const orc_robot *r, const orc_cfg *c, int B, int nthreads, const double *x0,
                          const double *ff, const double *caug, const double *xref, const double *noise, double *u,
                          double *x, double *cost_hist, double *e_u_hist, int *iters, int *status,
                          int *qp_stats /* 2*B: total GI iterations, max active-set size; or NULL */) {
  const int nj = r->nj, ns = 2 * nj, H = c->H, n = nj * H, N = ns * H;
  double *J0 = NULL;
  if (c->solver == 0) {
    J0 = (double *)malloc(sizeof(double) * n * n);
    if (orc_chol_J0(n, c->QQ, J0)) {
      for (int b = 0; b < B; ++b) {
        status[b] = ORC_NUMERICAL;
        iters[b] = 0;
      }
      free(J0);
      return;
    }
  }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    for (int k = 0; k < c->max_outer; ++k) cost_hist[(size_t)b * c->max_outer + k] = NAN;
    if (e_u_hist)
      for (int k = 0; k < c->max_outer; ++k) e_u_hist[(size_t)b * c->max_outer + k] = NAN;
    status[b] = orc_cfs_solve(r, c, J0, x0 + (size_t)b * ns, ff + (size_t)b * n, caug[b], xref + (size_t)b * N,
                              noise ? noise + (size_t)b * n * c->max_outer : NULL, u + (size_t)b * n,
                              x + (size_t)b * N, cost_hist + (size_t)b * c->max_outer,
                              e_u_hist ? e_u_hist + (size_t)b * c->max_outer : NULL, &iters[b],
                              qp_stats ? qp_stats + 2 * b : NULL);
  }
  if (J0) free(J0);
}

/* ------------------------------------------------------------------------------------------
 * CHOMP_FANUC (Lib/CHOMP_FANUC.m), the gradient-descent baseline planner.
 *   dm_f (:115-134): per-link distance minus obs.D with DH(i,1)=theta(i) and NO joint-2 offset, also for
 *     the 200i (unlike dist_link_200i.m:8) -- restated as written; touch rule as everywhere else.
 *   dcostObs_f (:137-165): per waypoint i and obstacle j: linkid = argmin dm_f, gradient of
 *     dist_link_*(linkid) by derivest per joint (:151,:156), chained through
 *     Baug((i-1)*njoint+1 : i*njoint, :)  (:153,:158).  NOTE (faithful quirk): the row stride is njoint,
 *     not nstate = 2*njoint, so waypoint i (1-based) picks the THETA rows of step (i+1)/2 when i is odd and
 *     the OMEGA rows of step i/2 when i is even.
 *   CHOMP_update_arm (:73-83): u <- u - alpha*3*(QQ*u + ff + 2000*dcostObs), then the roll-out.
 *   fobs_m (:91-112): sum over waypoints, obstacles and ALL links of the CHOMP potential of dm_f.
 *   optimizer (:54-69): eval.x_ / eval.x_old are never updated, so stop_outer only ends at iter_O >
 *     MAX_O_ITER: exactly max_outer updates.  cost_all(k) = get_cost(u_k) + fobs_m(x_k) (:63).
 * ---------------------------------------------------------------------------------------- */
static void chomp_dm(const orc_robot *r, const double *theta, const double *obs6, double D, double *d /* nj */,
                     int *touched) {
  double pos[ORC_MAXL * 6];
  if (r->kind == ORC_2L) {
    cap_pos_2l(r, theta, pos);
  } else {
    double DH[ORC_MAXL][4];
    memcpy(DH, r->DH, sizeof(DH));
    for (int i = 0; i < r->nj; ++i) DH[i][0] = theta[i]; /* :119-121, no offset */
    cap_pos_dh(DH, r->nj, r->base, r->cap, pos);
  }
  for (int i = 0; i < r->nj; ++i) d[i] = link_dist(pos + 6 * i, obs6, touched) - D; /* :127-133 */
}

static double chomp_potential(double d, double eps) { /* :100-106 */
  if (d < 0) return -d + 0.5 * eps;
  if (d <= eps) return (1.0 / (2.0 * eps)) * (d - eps) * (d - eps);
  return 0.0;
}

int orc_chomp_solve(const orc_robot *r, const orc_cfg *c, const double *D, const double *eps, const double *x0,
                    const double *ff, double caug, const double *xref, const double *u_init, double *u, double *x,
                    double *cost_hist, double *e_u_hist, int *iters, int *touched) {
  const int nj = r->nj, H = c->H, n = nj * H, N = 2 * nj * H;
  const double dt = r->dt;
  double *g = (double *)malloc(sizeof(double) * 2 * n), *uo = g + n;
  for (int k = 0; k < n; ++k) u[k] = u_init[k];
  for (int k = 0; k < N; ++k) x[k] = xref[k];
  int it = 0;
  for (it = 1; it <= c->max_outer; ++it) {
    for (int k = 0; k < n; ++k) uo[k] = u[k];
    /* dcostObs_f at the current x_ */
    for (int k = 0; k < n; ++k) g[k] = 0.0;
    for (int i = 1; i <= H; ++i) {
      const double *theta = x + (size_t)(i - 1) * 2 * nj;
      for (int j = 0; j < c->nobs; ++j) {
        const double *o6 = c->obs + (size_t)ORC_OBS_STRIDE * j;
        double d[ORC_MAXL];
        chomp_dm(r, theta, o6, D[j], d, touched);
        int lid = 1;
        for (int s = 1; s < nj; ++s)
          if (d[s] < d[lid - 1]) lid = s + 1; /* [dis, linkid] = min(Dfx): first minimum */
        const double dv = d[lid - 1];
        double w;
        if (dv < 0) w = -1.0;
        else if (dv <= eps[j]) w = (1.0 / eps[j]) * (dv - eps[j]);
        else continue;
        double dD[ORC_MAXL];
        orc_derivest_grad(r, theta, o6, lid, dD, touched);
        /* rows (i-1)*nj+1 .. i*nj of Baug: theta rows of step (i+1)/2 (i odd) or omega rows of step i/2 (i even) */
        const int step = (i + 1) / 2, odd = i & 1;
        for (int jj = 1; jj <= step; ++jj)
          for (int k = 0; k < nj; ++k) {
            const double bcoef = odd ? (0.5 * dt * dt + ((step - jj) * dt) * dt) : dt;
            g[(jj - 1) * nj + k] += w * dD[k] * bcoef;
          }
      }
    }
    /* u <- u - alpha*3*(QQ*u + ff + 2000*dcostObs)   (:75) */
    for (int cidx = 0; cidx < n; ++cidx) {
      double qu = 0;
      for (int k = 0; k < n; ++k) qu += c->QQ[cidx + (size_t)n * k] * uo[k];
      u[cidx] = uo[cidx] - (c->alpha * 3) * ((qu + ff[cidx]) + 2000 * g[cidx]);
    }
    rollout(nj, H, dt, x0, u, x);
    /* cost_new = get_cost(u) + fobs_m()   (:63) */
    double fobs = 0;
    for (int i = 1; i <= H; ++i)
      for (int j = 0; j < c->nobs; ++j) {
        double d[ORC_MAXL];
        chomp_dm(r, x + (size_t)(i - 1) * 2 * nj, c->obs + (size_t)ORC_OBS_STRIDE * j, D[j], d, touched);
        for (int s = 0; s < nj; ++s) fobs += chomp_potential(d[s], eps[j]);
      }
    cost_hist[it - 1] = get_cost(n, c->QQ, ff, caug, u) + fobs;
    if (e_u_hist) {
      double e2 = 0;
      for (int k = 0; k < n; ++k) e2 += (uo[k] - u[k]) * (uo[k] - u[k]);
      e_u_hist[it - 1] = sqrt(e2);
    }
  }
  *iters = c->max_outer;
  free(g);
  return ORC_MAX_ITER;
}

void orc_chomp_solve_batch(const orc_robot *r, const orc_cfg *c, const double *D, const double *eps, int B, int nthreads,
                           const double *x0, const double *ff, const double *caug, const double *xref,
                           const double *u_init, double *u, double *x, double *cost_hist, double *e_u_hist, int *iters,
                           int *status) {
  const int nj = r->nj, n = nj * c->H, N = 2 * n;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  (void)nthreads;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    int touched = 0;
    status[b] = orc_chomp_solve(r, c, D, eps, x0 + (size_t)b * 2 * nj, ff + (size_t)b * n, caug[b], xref + (size_t)b * N,
                                u_init + (size_t)b * n, u + (size_t)b * n, x + (size_t)b * N,
                                cost_hist + (size_t)b * c->max_outer,
                                e_u_hist ? e_u_hist + (size_t)b * c->max_outer : NULL, &iters[b], &touched);
    if (touched) status[b] |= 0x100;
  }
}

/* ------------------------------------------------------------------------------------------
 * RRT_FANUC.feasible (Lib/RRT_FANUC.m:146-181): infeasible if any link distance < obs{j}.D (:172).
 * Returns 1 feasible / 0 not; *dmin = min over all (obstacle, link) distances (no early break, so that
 * the value is order-independent; the reference breaks out of the link loop, which only saves work).
 * ---------------------------------------------------------------------------------------- */
int orc_rrt_feasible(const orc_robot *r, const double *theta, int nobs, const double *obs, const double *D,
                     double *dmin, int *touched) {
  double pos[ORC_MAXL * 6];
  int feas = 1;
  double dm = INFINITY;
  for (int j = 0; j < nobs; ++j) {
    orc_cap_pos(r, theta, pos);
    for (int i = 0; i < r->nj; ++i) {
      const double dis = link_dist(pos + 6 * i, obs + ORC_OBS_STRIDE * j, touched);
      if (dis < dm) dm = dis;
      if (dis < D[j]) feas = 0;
    }
  }
  if (dmin) *dmin = dm;
  return feas;
}

/* RRT_FANUC.getRandNode nearest scan (Lib/RRT_FANUC.m:116-127): argmin ||(node-sample).*ratial||, first wins */
int orc_rrt_nearest(int nj, int nnodes, const double *nodes, const double *sample, const double *ratial,
                    double *dists) {
  int parent = 0;
  double best = 0;
  for (int i = 0; i < nnodes; ++i) {
    double s = 0;
    for (int k = 0; k < nj; ++k) {
      const double v = (nodes[i * nj + k] - sample[k]) * ratial[k];
      s += v * v;
    }
    const double di = sqrt(s);
    if (dists) dists[i] = di;
    if (i == 0 || di < best) {
      best = di;
      parent = i;
    }
  }
  return parent;
}

/* ------------------------------------------------------------------------------------------
 * RRT_FANUC.find_route (Lib/RRT_FANUC.m:63-93) with getNode/getRandNode (:97-131), arrangeNode (:134-142),
 * addNode (:184-190) and goal_reached (:193-207).  MATLAB's rand stream is an INPUT: rnd[] is consumed in the
 * reference's order -- pp = rand (:108), then rand(nstate,1) if pp < bi (:111).
 *   star = 1: 'RRT*' (arrangeNode after every addNode), 0: 'RRT'
 *   nodes (nj x cap), parent (1-based column index, -1 for the root), total_dis: the tree; cap >= max_iter + 1
 *   route (nj x cap): the root-to-last-node path (:86-91); returns its length (size(route,2) = routeL of
 *   s_Parallel_rrt.m:21); *fail: node_num > MAX_ITER (:201-205); *rnd_used: numbers consumed.
 * Returns -1 when the random stream is exhausted before the search ends.
 * ---------------------------------------------------------------------------------------- */
int orc_rrt_find_route(const orc_robot *r, int nobs, const double *obs, const double *D, const double *x0,
                       const double *goal, const double *region_g, const double *region_s, const double *sample_off,
                       const double *goal_th, const double *ratial, double bi, int max_iter, int star,
                       const double *rnd, int nrnd, int cap, double *nodes, int *parent, double *total_dis,
                       double *route, int *n_nodes, int *fail, int *rnd_used) {
  const int nj = r->nj;
  double *to_dis = (double *)malloc(sizeof(double) * (size_t)cap);
  double newn[ORC_MAXL], sample[ORC_MAXL];
  int node_num = 1, cur = 0, par = 0 /* [] in MATLAB: the loop below never ran */, reached = 0, touched = 0;
  *fail = 0;
  for (int k = 0; k < nj; ++k) nodes[k] = newn[k] = x0[k];
  parent[0] = -1;
  total_dis[0] = 0.0;
  for (;;) {
    /* goal_reached (:193-207): both comparisons must hold for every joint */
    int in = 1;
    for (int k = 0; k < nj; ++k)
      if (!((goal[k] - region_g[k]) < newn[k] && newn[k] < (goal[k] + region_g[k]))) in = 0;
    reached = in;
    if (node_num > max_iter) {
      *fail = 1;
      reached = 1;
    }
    if (reached) break;
    /* getNode (:97-104) */
    for (;;) {
      if (cur >= nrnd) { free(to_dis); return -1; }
      const double pp = rnd[cur++];
      if (pp < bi) {
        if (cur + nj > nrnd) { free(to_dis); return -1; }
        for (int k = 0; k < nj; ++k) sample[k] = (rnd[cur + k] - 0.5) * region_s[k] * 2 + sample_off[k];
        cur += nj;
      } else {
        for (int k = 0; k < nj; ++k) sample[k] = goal_th[k];
      }
      par = 1 + orc_rrt_nearest(nj, node_num, nodes, sample, ratial, to_dis);
      const double *pn = nodes + (size_t)(par - 1) * nj;
      double ss = 0.0;
      for (int k = 0; k < nj; ++k) ss += (pn[k] - sample[k]) * (pn[k] - sample[k]);
      const double nrm = sqrt(ss);
      for (int k = 0; k < nj; ++k) newn[k] = pn[k] + (sample[k] - pn[k]) * 0.1 / nrm; /* :129 */
      if (orc_rrt_feasible(r, newn, nobs, obs, D, NULL, &touched)) break;
    }
    /* addNode (:184-190) */
    for (int k = 0; k < nj; ++k) nodes[(size_t)node_num * nj + k] = newn[k];
    parent[node_num] = par;
    total_dis[node_num] = total_dis[par - 1] + to_dis[par - 1];
    ++node_num;
    /* arrangeNode (:134-142): nodes closer than 0.2 to the SAMPLE are re-parented to the new node when that is shorter */
    if (star)
      for (int i = 0; i < node_num - 1; ++i)
        if (to_dis[i] < 0.2 && total_dis[i] > (total_dis[node_num - 1] + to_dis[i])) {
          parent[i] = node_num;
          total_dis[i] = total_dis[node_num - 1] + to_dis[i];
        }
  }
  free(to_dis);
  /* route (:86-91) */
  int len = 1, p = par, guard = 0;
  while (p > 0 && guard++ <= node_num) { ++len; p = parent[p - 1]; }
  int pos = len - 1;
  for (int k = 0; k < nj; ++k) route[(size_t)pos * nj + k] = newn[k];
  p = par; guard = 0;
  while (p > 0 && guard++ <= node_num) {
    --pos;
    for (int k = 0; k < nj; ++k) route[(size_t)pos * nj + k] = nodes[(size_t)(p - 1) * nj + k];
    p = parent[p - 1];
  }
  *n_nodes = node_num;
  *rnd_used = cur;
  return len;
}
