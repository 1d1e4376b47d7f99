/*
 * cfs_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C FP64 restatement of the MATLAB reference JessicaLeu-code/MotionPlanning_5D_m
 * for the Convex-Feasible-Set hot path.  Every function cites the reference file:line it
 * follows (paths relative to the reference root).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (libcfs_b200.so) never links or calls it.
 *
 * PARITY PINNING.  Geometry/calculus parts are pinned against the reference's own known
 * answers: distLinSeg doc example (Lib/functions/distLinSeg.m:15-18) and the DERIVEST
 * header/demo values (DERIVESTsuite/DERIVESTsuite/derivest.m:163-174, demo/derivest_demo.m).
 * The QP step is "PARITY UNPINNED": the reference calls MathWorks' closed-source quadprog
 * (Lib/CFS_FANUC.m:85, Lib/PSGCFS_FANUC.m:120), which is not under the reference tree and
 * has no pinned version; MATLAB/Octave are not installed here.  The QPs are strictly convex
 * so the optimum is unique; this oracle solves them with a dense Goldfarb-Idnani dual
 * active-set method and reports the KKT residual so tests can assert optimality.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp)
 */
#ifndef CFS_ORACLE_H
#define CFS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXL 6

enum { ORC_M16IB = 0, ORC_M200I = 1, ORC_2L = 2 };

/* status codes shared with include/cfs_b200.h */
enum {
  ORC_OK_CONVERGED = 0, /* ||x_-x_old|| < epsilon_O            (EVAL.m:64-67) */
  ORC_MAX_ITER = 1,     /* iter_O > MAX_O_ITER                  (EVAL.m:69-72) */
  ORC_QP_INFEASIBLE = 2,/* quadprog would return [] and the rollout would throw (CFS_FANUC.m:85-92) */
  ORC_NUMERICAL = 3,
  ORC_FLAG_TOUCH = 0x100 /* |dis|<1e-4 branch taken (dist_arm_3D_Heu_2.m:22-24): on M16iB the reference throws */
};

typedef struct {
  int kind;                     /* ORC_M16IB / ORC_M200I / ORC_2L                            */
  int nj;                       /* joints (= links used): 5, 5, 2                            */
  double DH[ORC_MAXL][4];       /* theta d a alpha           robotproperty2.m:24-29,68-73    */
  double base[3];               /*                           robotproperty2.m:53-54,98-99    */
  double cap[ORC_MAXL][2][3];   /* cap{i}.p(:,k)             robotproperty2.m:36-52,77-94    */
  double T2L[3][3];             /* robot.T(:,c)  (2L only)   robotproperty2.m:117-119        */
  double dt;                    /* delta_t                                                   */
} orc_robot;

typedef struct {
  int H;                /* horizon                                                            */
  int nobs;             /* obstacles                                                          */
  const double *obs;    /* ORC_OBS_STRIDE*nobs: obs{j}.l(:,1), obs{j}.l(:,2), kind (0 capsule axis, 1 box) */
  const double *margin; /* nobs: obs{j}.epsilon (CFS_FANUC.m:117) or obs{j}.D (PSGCFS:158)    */
  const double *QQ;     /* n*n (symmetric)                                                    */
  const double *lim;    /* nj   velocity limit, NULL = no velocity rows (M16iB/main_CFS.m)    */
  const double *max_input; /* n bounds, NULL = no bounds (PSGCFS projection has none)         */
  double eps_outer;     /* sys_info.epsilon_O                                                 */
  int max_outer;        /* sys_info.MAX_O_ITER                                                */
  int solver;           /* 0 = CFS_FANUC, 1 = PSGCFS_FANUC                                    */
  int grad;             /* 0 = num_jac(dist_arm), 1 = derivest(dist_link(linkid))             */
  double alpha;         /* PSG step (sys_info.alpha)                                          */
} orc_cfg;

#define ORC_OBS_STRIDE 7 /* doubles per obstacle record: l(:,1), l(:,2), kind */
void orc_robot_init(orc_robot *r, int kind);
/* N3 extension: distance between a segment and a solid axis-aligned box, closest point of the segment in point[3] */
double orc_dist_seg_box(const double *ps, const double *pe, const double *lo, const double *hi, double *point);

/* geometry */
void   orc_cap_pos(const orc_robot *r, const double *theta, double *pos /* nj*2*3 */);
double orc_dist_lin_seg(const double *p1s, const double *p1e, const double *p2s, const double *p2e,
                        int dim, double *points /* 2*dim or NULL */);
double orc_dist_arm(const orc_robot *r, const double *theta, const double *obs6, int *linkid, int *touched);
double orc_dist_link(const orc_robot *r, const double *theta, const double *obs6, int linkid, int *touched);
void   orc_num_jac(const orc_robot *r, const double *theta, const double *obs6, double *grad, int *touched);

/* DERIVEST (derivest.m defaults + 'Vectorized','no') */
typedef double (*orc_fun1)(double x, void *ctx);
void orc_derivest(orc_fun1 f, void *ctx, double x0, double *der, double *errest, double *finaldelta);
void orc_derivest_grad(const orc_robot *r, const double *theta, const double *obs6, int linkid,
                       double *grad, int *touched);
/* named test integrands for the DERIVEST known answers: 0 exp,1 sin,2 sinh,3 log,4 x^3+x^4-ish unused */
void orc_derivest_named(int which, double x0, double *der, double *errest, double *finaldelta);

/* cost builder, main_FANUC.m:64-103 */
void orc_build_cost(int nj, int H, double dt, const double *Q /*2nj x 2nj cm*/, const double *Rblk /*nj x nj cm*/,
                    double r_scale, double stage_w, double term_w,
                    double *Aaug /*2njH x 2nj cm*/, double *Baug /*2njH x njH cm*/, double *Qaug_diagblocks /*unused may be NULL*/,
                    double *QQ /* n x n cm */);
void orc_build_ff(int nj, int H, const double *Q, double stage_w, double term_w, const double *Aaug,
                  const double *Baug, const double *x0, const double *gaug /*2njH*/, double *ff, double *caug);

/* constraints, CFS_FANUC.m:101-135 (row-major Ainq m x n, m = nobs*H*(1+2nj) or nobs*H) */
int orc_get_con(const orc_robot *r, const orc_cfg *c, const double *x0, const double *xcur, const double *u,
                double *Ainq, double *binq, double *dist /*H*nobs*/, int *linkid /*H*nobs*/, double *grad /*nj*H*nobs*/,
                int *touched);

/* strictly convex QP  min 1/2 x'Gx + a'x  s.t. C x <= d   (C row-major m x n); Goldfarb-Idnani.
 * J0 = L^{-T} (n x n col-major) with G = L L'.  returns 0 ok, 2 infeasible, 3 numerical. */
int orc_qp_gi(int n, int m, const double *J0, const double *xunc, const double *C, const double *d,
              double *x, double *lam /* m or NULL */, int *iters);
int orc_qp_gi2(int n, int m, const double *J0, const double *xunc, const double *C, const double *d,
               double *x, double *lam, int *iters, int *qmax);
int orc_chol_J0(int n, const double *G, double *J0);
double orc_kkt_residual(int n, int m, const double *G, const double *a, const double *C, const double *d,
                        const double *x, const double *lam);

/* full solves */
int orc_cfs_solve(const orc_robot *r, const orc_cfg *c, const double *J0 /* may be NULL */,
                  const double *x0, const double *ff, double caug, const double *xref,
                  const double *noise /* n*max_outer or NULL */,
                  double *u, double *x, double *cost_hist, double *e_u_hist /* may be NULL */,
                  int *iters, int *qp_stats /* [2]: GI iterations, max active set; may be NULL */);
void orc_cfs_solve_batch(const orc_robot *r, const orc_cfg *c, int B, int nthreads,
                         const double *x0, const double *ff, const double *caug, const double *xref,
                         const double *noise,
                         double *u, double *x, double *cost_hist, double *e_u_hist, int *iters, int *status);

void orc_cfs_solve_batch2(const orc_robot *r, const orc_cfg *c, int B, int nthreads,
                          const double *x0, const double *ff, const double *caug, const double *xref,
                          const double *noise,
                          double *u, double *x, double *cost_hist, double *e_u_hist, int *iters, int *status,
                          int *qp_stats /* 2*B or NULL */);

/* CHOMP_FANUC.optimizer (Lib/CHOMP_FANUC.m:54-165): exactly c->max_outer gradient steps from u_init / xref.
 * D, eps: obs{j}.D and obs{j}.epsilon (c->margin is not used).  Returns ORC_MAX_ITER. */
int orc_chomp_solve(const orc_robot *r, const orc_cfg *c, const double *D, const double *eps, const double *x0,
                    const double *ff, double caug, const double *xref, const double *u_init, double *u, double *x,
                    double *cost_hist, double *e_u_hist, int *iters, int *touched);
void orc_chomp_solve_batch(const orc_robot *r, const orc_cfg *c, const double *D, const double *eps, int B, int nthreads,
                           const double *x0, const double *ff, const double *caug, const double *xref,
                           const double *u_init, double *u, double *x, double *cost_hist, double *e_u_hist, int *iters,
                           int *status);

/* RRT_FANUC.feasible (RRT_FANUC.m:146-181) and nearest/steer (RRT_FANUC.m:116-129) */
int  orc_rrt_feasible(const orc_robot *r, const double *theta, int nobs, const double *obs, const double *D,
                      double *dmin, int *touched);
int  orc_rrt_nearest(int nj, int nnodes, const double *nodes /* nj x nnodes */, const double *sample,
                     const double *ratial, double *dists /* nnodes or NULL */);

/* RRT_FANUC.find_route (RRT_FANUC.m:63-207) driven by a caller-supplied uniform random stream; see cfs_oracle.c */
int  orc_rrt_find_route(const orc_robot *r, int nobs, const double *obs, const double *D, const double *x0,
                        const double *goal, const double *region_g, const double *region_s, const double *sample_off,
                        const double *goal_th, const double *ratial, double bi, int max_iter, int star,
                        const double *rnd, int nrnd, int cap, double *nodes, int *parent, double *total_dis,
                        double *route, int *n_nodes, int *fail, int *rnd_used);

#ifdef __cplusplus
}
#endif
#endif
