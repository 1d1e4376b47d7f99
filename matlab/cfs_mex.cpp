// cfs_mex.cpp -- MEX gateway between the reference's MATLAB host code and libcfs_b200.so (include/cfs_b200.h).
//
//   [u, x_, cost_all, e_u_all, iters, status, qp_steps] = cfs_mex(solver, grad, ROBOT, obs, sys_info [, noise])
//
//     solver   'CFS' | 'PSGCFS'                       (Lib/CFS_FANUC.m / Lib/PSGCFS_FANUC.m)
//     grad     'num_jac' | 'derivest'                 (Lib/functions/num_jac.m / DERIVESTsuite derivest.m on dist_link_*)
//     ROBOT    'M16iB' | 'M200i' | '2L'               (CFS_FANUC.m:49-54)
//     obs      cell of structs with fields l (3x2), D, epsilon                    (main_FANUC.m:56-60)
//     sys_info struct exactly as the mains build it                              (main_FANUC.m:106-127)
//              batched use: sys_info.xR (nstate x B), .ff (n x B), .caug (1 x B), .x_ (nstate*H x B)
//     noise    n x MAX_O_ITER x B normrnd(0,0.1) draws for PSGCFS (PSGCFS_FANUC.m:109); omitted = zeros
//
// Build (on a machine with MATLAB; this container has no mex.h, so the file is shipped uncompiled):
//   mex -I<repo>/include cfs_mex.cpp -L<repo>/motionplanning_5d_m_b200 -lcfs_b200
// The gateway owns one cfs_ctx per MATLAB process (parfor workers each load their own copy, s_Parallel_rrt.m:16).
#include <cstring>
#include <string>
#include <vector>

#include "cfs_b200.h"
#include "mex.h"

static cfs_ctx *g_ctx = nullptr;

static void at_exit() {
  if (g_ctx) cfs_destroy(g_ctx);
  g_ctx = nullptr;
}

static void check(int rc, const char *what) {
  if (rc != 0) mexErrMsgIdAndTxt("cfs:cuda", "%s failed (%d): %s", what, rc, cfs_last_error(g_ctx));
}

static const mxArray *field(const mxArray *s, const char *name, bool required = true) {
  const mxArray *f = mxGetField(s, 0, name);
  if (!f && required) mexErrMsgIdAndTxt("cfs:arg", "missing field '%s'", name);
  return f;
}

static std::string str(const mxArray *a) {
  char buf[64];
  if (mxGetString(a, buf, sizeof(buf))) mexErrMsgIdAndTxt("cfs:arg", "expected a short char array");
  return buf;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 5) mexErrMsgIdAndTxt("cfs:arg", "usage: cfs_mex(solver, grad, ROBOT, obs, sys_info [, noise])");
  if (!g_ctx) {
    cfs_ctx *c = nullptr;
    if (cfs_create(&c, 0) != 0) mexErrMsgIdAndTxt("cfs:cuda", "%s", cfs_last_error(nullptr));  // no CPU fallback
    g_ctx = c;
    mexAtExit(at_exit);
  }
  const std::string solver = str(prhs[0]), grad = str(prhs[1]), robot_name = str(prhs[2]);
  const mxArray *obs = prhs[3], *si = prhs[4];
  const int kind = robot_name == "M16iB" ? CFS_ROBOT_M16IB : robot_name == "M200i" ? CFS_ROBOT_M200I : CFS_ROBOT_2L;
  const int H = (int)mxGetScalar(field(si, "H")), nj = (int)mxGetScalar(field(si, "njoint"));
  const int n = H * nj;

  // ---- robot (robotproperty2.m) -------------------------------------------------------------------------------
  const mxArray *rb = field(si, "robot");
  const mxArray *DH = field(rb, "DH"), *cap = field(rb, "cap");
  std::vector<double> cap_p(6 * nj);
  for (int i = 0; i < nj; ++i) std::memcpy(&cap_p[6 * i], mxGetPr(mxGetField(mxGetCell(cap, i), 0, "p")), 6 * sizeof(double));
  const mxArray *T = mxGetField(rb, 0, "T");
  check(cfs_set_robot(g_ctx, kind, mxGetPr(DH), (int)mxGetM(DH), mxGetPr(field(rb, "base")), cap_p.data(), nj,
                      T ? mxGetPr(T) : nullptr, mxGetScalar(field(rb, "delta_t"))), "cfs_set_robot");
  // ---- obstacles (main_FANUC.m:56-60) ----------------------------------------------------------------------------
  const int nobs = (int)mxGetNumberOfElements(obs);
  std::vector<double> seg(6 * nobs), D(nobs), eps(nobs);
  for (int j = 0; j < nobs; ++j) {
    const mxArray *o = mxGetCell(obs, j);
    std::memcpy(&seg[6 * j], mxGetPr(field(o, "l")), 6 * sizeof(double));
    D[j] = mxGetScalar(field(o, "D"));
    eps[j] = mxGetScalar(field(o, "epsilon"));
  }
  check(cfs_set_obstacles(g_ctx, seg.data(), D.data(), eps.data(), nobs), "cfs_set_obstacles");
  // ---- cost (main_FANUC.m:106-127); PSGCFS projects without bounds (PSGCFS_FANUC.m:120) -----------------------------
  const bool psg = solver == "PSGCFS";
  const mxArray *lim = field(si, "lim", false), *mi = field(si, "MAX_input", false);
  check(cfs_set_cost(g_ctx, H, mxGetPr(field(si, "QQ")), lim ? mxGetPr(lim) : nullptr, (mi && !psg) ? mxGetPr(mi) : nullptr),
        "cfs_set_cost");
  // ---- the batch ----------------------------------------------------------------------------------------------------
  const mxArray *xR = field(si, "xR"), *ff = field(si, "ff"), *caug = field(si, "caug"), *xref = field(si, "x_");
  const int B = (int)mxGetN(ff);
  const int K = (int)mxGetScalar(field(si, "MAX_O_ITER"));
  std::vector<double> x0((size_t)2 * nj * B);
  const size_t ldxr = mxGetM(xR) * (mxGetN(xR) / (size_t)B);  // sys_info.xR may carry later roll-out columns (B = 1)
  for (int b = 0; b < B; ++b) std::memcpy(&x0[(size_t)2 * nj * b], mxGetPr(xR) + ldxr * b, 2 * nj * sizeof(double));
  const double *noise = (nrhs > 5 && !mxIsEmpty(prhs[5])) ? mxGetPr(prhs[5]) : nullptr;
  const mxArray *alpha = field(si, "alpha", false);
  plhs[0] = mxCreateDoubleMatrix(n, B, mxREAL);
  mxArray *x = mxCreateDoubleMatrix(2 * n, B, mxREAL), *cost = mxCreateDoubleMatrix(K, B, mxREAL),
          *eu = mxCreateDoubleMatrix(K, B, mxREAL);
  mxArray *it = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL), *st = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
  check(cfs_solve_batch(g_ctx, B, psg ? CFS_SOLVER_PSGCFS : CFS_SOLVER_CFS, grad == "derivest" ? CFS_GRAD_DERIVEST : CFS_GRAD_NUMJAC,
                        x0.data(), mxGetPr(ff), mxGetPr(caug), mxGetPr(xref), noise, mxGetScalar(field(si, "epsilon_O")), K,
                        alpha ? mxGetScalar(alpha) : 0.0, mxGetPr(plhs[0]), mxGetPr(x), mxGetPr(cost), mxGetPr(eu),
                        (int *)mxGetData(it), (int *)mxGetData(st)), "cfs_solve_batch");
  if (nlhs > 1) plhs[1] = x;
  if (nlhs > 2) plhs[2] = cost;
  if (nlhs > 3) plhs[3] = eu;
  if (nlhs > 4) plhs[4] = it;
  if (nlhs > 5) plhs[5] = st;
  if (nlhs > 6) {
    cfs_stats s;
    cfs_get_stats(g_ctx, &s);
    plhs[6] = mxCreateDoubleScalar((double)s.qp_steps);
  }
}
