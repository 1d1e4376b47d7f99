// harness.cpp -- implements the mex.h stub and a C entry point that builds the MATLAB-side arguments of
//   [u, x_, cost_all, e_u_all, iters, status, qp_steps] = cfs_mex(solver, grad, ROBOT, obs, sys_info [, noise])
// from plain arrays (ctypes), calls the gateway's mexFunction and copies the outputs back.  Test infrastructure only.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "mex.h"

static std::vector<mxArray *> g_all;
static void (*g_atexit)(void) = nullptr;
static mxArray *mk(mxClassID cls, size_t m, size_t n) {
  mxArray *a = new mxArray;
  a->cls = cls; a->m = m; a->n = n;
  if (cls == mxDOUBLE_CLASS) a->d.assign(m * n, 0.0);
  if (cls == mxINT32_CLASS) a->i32.assign(m * n, 0);
  g_all.push_back(a);
  return a;
}
double *mxGetPr(const mxArray *a) { return const_cast<double *>(a->d.data()); }
void *mxGetData(const mxArray *a) {
  return a->cls == mxINT32_CLASS ? (void *)const_cast<int *>(a->i32.data()) : (void *)const_cast<double *>(a->d.data());
}
double mxGetScalar(const mxArray *a) { return a->cls == mxINT32_CLASS ? (double)a->i32.at(0) : a->d.at(0); }
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->cls == mxCELL_CLASS ? a->cells.size() : a->m * a->n; }
bool mxIsEmpty(const mxArray *a) { return mxGetNumberOfElements(a) == 0; }
bool mxIsDouble(const mxArray *a) { return a && a->cls == mxDOUBLE_CLASS; }
bool mxIsStruct(const mxArray *a) { return a && a->cls == mxSTRUCT_CLASS; }
bool mxIsCell(const mxArray *a) { return a && a->cls == mxCELL_CLASS; }
bool mxIsChar(const mxArray *a) { return a && a->cls == mxCHAR_CLASS; }
mxArray *mxGetField(const mxArray *s, size_t index, const char *name) {
  if (!s || s->cls != mxSTRUCT_CLASS || index != 0) return nullptr;
  auto it = s->fields.find(name);
  return it == s->fields.end() ? nullptr : it->second;
}
mxArray *mxGetCell(const mxArray *c, size_t index) { return (c && index < c->cells.size()) ? c->cells[index] : nullptr; }
int mxGetString(const mxArray *a, char *buf, size_t buflen) {
  if (!a || a->cls != mxCHAR_CLASS || a->str.size() + 1 > buflen) return 1;
  std::strcpy(buf, a->str.c_str());
  return 0;
}
mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity) { return mk(mxDOUBLE_CLASS, m, n); }
mxArray *mxCreateDoubleScalar(double v) {
  mxArray *a = mk(mxDOUBLE_CLASS, 1, 1);
  a->d[0] = v;
  return a;
}
mxArray *mxCreateNumericMatrix(size_t m, size_t n, mxClassID cls, mxComplexity) { return mk(cls, m, n); }
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw mex_error{id, buf};
}
int mexAtExit(void (*fn)(void)) {
  g_atexit = fn;
  return 0;
}

static mxArray *dbl(size_t m, size_t n, const double *src) {
  mxArray *a = mk(mxDOUBLE_CLASS, m, n);
  if (src) std::memcpy(a->d.data(), src, sizeof(double) * m * n);
  return a;
}
static mxArray *chr(const char *s) {
  mxArray *a = mk(mxCHAR_CLASS, 1, std::strlen(s));
  a->str = s;
  return a;
}
static mxArray *strct() { return mk(mxSTRUCT_CLASS, 1, 1); }

static char g_err[1200];
extern "C" const char *mexh_last_error() { return g_err; }
extern "C" void mexh_shutdown() {
  if (g_atexit) g_atexit();
  g_atexit = nullptr;
}

static mxArray *robot_struct(int nj, int dh_rows, const double *DH, const double *base, const double *cap_p, const double *T, double dt) {
  mxArray *rb = strct();
  rb->fields["DH"] = dbl(dh_rows, 4, DH);
  rb->fields["base"] = dbl(3, 1, base);
  rb->fields["delta_t"] = dbl(1, 1, &dt);
  if (T) rb->fields["T"] = dbl(3, 3, T);
  mxArray *cap = mk(mxCELL_CLASS, 1, nj);
  for (int i = 0; i < nj; ++i) {
    mxArray *c = strct();
    c->fields["p"] = dbl(3, 2, cap_p + 6 * i);
    cap->cells.push_back(c);
  }
  rb->fields["cap"] = cap;
  return rb;
}
static mxArray *obs_cell(int nobs, const double *obs_l, const double *obs_D, const double *obs_eps) {
  mxArray *obs = mk(mxCELL_CLASS, 1, nobs);
  for (int j = 0; j < nobs; ++j) {
    mxArray *o = strct();
    o->fields["l"] = dbl(3, 2, obs_l + 6 * j);
    o->fields["D"] = dbl(1, 1, obs_D + j);
    o->fields["epsilon"] = dbl(1, 1, obs_eps + j);
    obs->cells.push_back(o);
  }
  return obs;
}
static int finish(int rc) {
  for (mxArray *a : g_all) delete a;
  g_all.clear();
  return rc;
}

// cfs_mex('device', id)
extern "C" int mexh_device(int id) {
  g_err[0] = 0;
  const double idd = id;
  const mxArray *prhs[2] = {chr("device"), dbl(1, 1, &idd)};
  mxArray *plhs[1] = {nullptr};
  int rc = 0;
  try {
    mexFunction(0, plhs, 2, prhs);
  } catch (const mex_error &e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", e.id.c_str(), e.msg.c_str());
    rc = 1;
  }
  return finish(rc);
}

// [routes, route_len, n_nodes, fail, rnd_used, nodes, parent, total_dis] = cfs_mex('rrt', ROBOT, SOLVER, obs, sys_info, goal,
//                                                                                 region_g, region_s, sample_off, rnd)
extern "C" int mexh_rrt(const char *robot_name, const char *solver, int nj, int dh_rows, const double *DH, const double *base,
                        const double *cap_p, double dt, int nobs, const double *obs_l, const double *obs_D, const double *obs_eps,
                        const double *x0, const double *goal, const double *goal_th, const double *ratial, const double *region_g,
                        const double *region_s, const double *sample_off, int nrnd, int S, const double *rnd, int max_iter,
                        double *routes /*nj x (max_iter+2) x S*/, int *route_len, int *n_nodes, int *fail, int *rnd_used,
                        int *parent /*(max_iter+2) x S*/, double *total_dis) {
  g_err[0] = 0;
  const int cap = max_iter + 2;
  mxArray *si = strct();
  const double njd = nj, bi = 0.5, mit = max_iter;
  si->fields["nstate"] = dbl(1, 1, &njd);
  si->fields["robot"] = robot_struct(nj, dh_rows, DH, base, cap_p, nullptr, dt);
  si->fields["x0"] = dbl(nj, 1, x0);
  si->fields["goal_th"] = dbl(nj, 1, goal_th);
  si->fields["ratial"] = dbl(nj, 1, ratial);
  const mxArray *prhs[12] = {chr("rrt"), chr(robot_name), chr(solver), obs_cell(nobs, obs_l, obs_D, obs_eps), si, dbl(nj, 1, goal),
                             dbl(nj, 1, region_g), dbl(nj, 1, region_s), dbl(nj, 1, sample_off), dbl(nrnd, S, rnd), dbl(1, 1, &bi),
                             dbl(1, 1, &mit)};
  mxArray *plhs[8] = {nullptr};
  int rc = 0;
  try {
    mexFunction(8, plhs, 12, prhs);
    std::memcpy(routes, plhs[0]->d.data(), sizeof(double) * nj * cap * S);
    std::memcpy(route_len, plhs[1]->i32.data(), sizeof(int) * S);
    std::memcpy(n_nodes, plhs[2]->i32.data(), sizeof(int) * S);
    std::memcpy(fail, plhs[3]->i32.data(), sizeof(int) * S);
    std::memcpy(rnd_used, plhs[4]->i32.data(), sizeof(int) * S);
    if (parent) std::memcpy(parent, plhs[6]->i32.data(), sizeof(int) * cap * S);
    if (total_dis) std::memcpy(total_dis, plhs[7]->d.data(), sizeof(double) * cap * S);
  } catch (const mex_error &e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", e.id.c_str(), e.msg.c_str());
    rc = 1;
  }
  return finish(rc);
}

// [u, x_, cost_all, e_u_all, iters, status] = cfs_mex('routes', ROBOT, obs, sys_info, routes, route_len, Q, Rblk, r_scale)
extern "C" int mexh_routes(const char *robot_name, int nj, int H, int B, int W, int dh_rows, const double *DH, const double *base,
                           const double *cap_p, double dt, int nobs, const double *obs_l, const double *obs_D, const double *obs_eps,
                           const double *lim, const double *max_input, double eps_outer, int max_outer, const double *routes,
                           const double *route_len /*B, as MATLAB doubles*/, const double *Q, const double *Rblk, double r_scale,
                           double *u, double *x, double *cost, double *eu, int *iters, int *status) {
  g_err[0] = 0;
  const size_t n = (size_t)H * nj;
  mxArray *si = strct();
  const double Hd = H, njd = nj, Kd = max_outer;
  si->fields["H"] = dbl(1, 1, &Hd);
  si->fields["njoint"] = dbl(1, 1, &njd);
  si->fields["robot"] = robot_struct(nj, dh_rows, DH, base, cap_p, nullptr, dt);
  if (lim) si->fields["lim"] = dbl(nj, 1, lim);
  if (max_input) si->fields["MAX_input"] = dbl(n, 1, max_input);
  si->fields["epsilon_O"] = dbl(1, 1, &eps_outer);
  si->fields["MAX_O_ITER"] = dbl(1, 1, &Kd);
  const mxArray *prhs[9] = {chr("routes"), chr(robot_name), obs_cell(nobs, obs_l, obs_D, obs_eps), si, dbl((size_t)nj * W, B, routes),
                            dbl(1, B, route_len), dbl(2 * nj, 2 * nj, Q), dbl(nj, nj, Rblk), dbl(1, 1, &r_scale)};
  mxArray *plhs[6] = {nullptr};
  int rc = 0;
  try {
    mexFunction(6, plhs, 9, prhs);
    std::memcpy(u, plhs[0]->d.data(), sizeof(double) * n * B);
    std::memcpy(x, plhs[1]->d.data(), sizeof(double) * 2 * n * B);
    std::memcpy(cost, plhs[2]->d.data(), sizeof(double) * max_outer * B);
    std::memcpy(eu, plhs[3]->d.data(), sizeof(double) * max_outer * B);
    std::memcpy(iters, plhs[4]->i32.data(), sizeof(int) * B);
    std::memcpy(status, plhs[5]->i32.data(), sizeof(int) * B);
  } catch (const mex_error &e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", e.id.c_str(), e.msg.c_str());
    rc = 1;
  }
  return finish(rc);
}

// All matrices column-major as MATLAB stores them.  T may be NULL (DH robots); lim / max_input / noise may be NULL.
// returns 0, or 1 with mexh_last_error() = "id: message" when the gateway raised a MATLAB error.
extern "C" int mexh_cfs(const char *solver, const char *grad, const char *robot_name, int nj, int H, int B, int dh_rows,
                        const double *DH, const double *base, const double *cap_p /*3x2 x nj*/, const double *T /*3x3*/,
                        double dt, int nobs, const double *obs_l /*3x2 x nobs*/, const double *obs_D, const double *obs_eps,
                        const double *QQ, const double *lim, const double *max_input, const double *xR /*2nj x B*/,
                        const double *ff /*n x B*/, const double *caug /*B*/, const double *xref /*2njH x B*/,
                        double eps_outer, int max_outer, double alpha, const double *noise /*n x K x B*/, double *u,
                        double *x, double *cost, double *eu, int *iters, int *status, double *qp_steps) {
  const size_t n = (size_t)H * nj;
  g_err[0] = 0;
  mxArray *rb = strct();
  rb->fields["DH"] = dbl(dh_rows, 4, DH);
  rb->fields["base"] = dbl(3, 1, base);
  rb->fields["delta_t"] = dbl(1, 1, &dt);
  if (T) rb->fields["T"] = dbl(3, 3, T);
  mxArray *cap = mk(mxCELL_CLASS, 1, nj);
  for (int i = 0; i < nj; ++i) {
    mxArray *c = strct();
    c->fields["p"] = dbl(3, 2, cap_p + 6 * i);
    cap->cells.push_back(c);
  }
  rb->fields["cap"] = cap;
  mxArray *obs = mk(mxCELL_CLASS, 1, nobs);
  for (int j = 0; j < nobs; ++j) {
    mxArray *o = strct();
    o->fields["l"] = dbl(3, 2, obs_l + 6 * j);
    o->fields["D"] = dbl(1, 1, obs_D + j);
    o->fields["epsilon"] = dbl(1, 1, obs_eps + j);
    obs->cells.push_back(o);
  }
  mxArray *si = strct();
  const double Hd = H, njd = nj, Kd = max_outer;
  si->fields["H"] = dbl(1, 1, &Hd);
  si->fields["njoint"] = dbl(1, 1, &njd);
  si->fields["robot"] = rb;
  si->fields["QQ"] = dbl(n, n, QQ);
  if (lim) si->fields["lim"] = dbl(nj, 1, lim);
  if (max_input) si->fields["MAX_input"] = dbl(n, 1, max_input);
  si->fields["xR"] = dbl(2 * nj, B, xR);
  si->fields["ff"] = dbl(n, B, ff);
  si->fields["caug"] = dbl(1, B, caug);
  si->fields["x_"] = dbl(2 * n, B, xref);
  si->fields["epsilon_O"] = dbl(1, 1, &eps_outer);
  si->fields["MAX_O_ITER"] = dbl(1, 1, &Kd);
  si->fields["alpha"] = dbl(1, 1, &alpha);
  const mxArray *prhs[6] = {chr(solver), chr(grad), chr(robot_name), obs, si, noise ? (std::strcmp(solver, "CHOMP") == 0 ? dbl(n, B, noise) /* uu */ : dbl(n * max_outer, B, noise)) : nullptr};
  mxArray *plhs[7] = {nullptr};
  int rc = 0;
  try {
    mexFunction(7, plhs, noise ? 6 : 5, prhs);
    std::memcpy(u, plhs[0]->d.data(), sizeof(double) * n * B);
    std::memcpy(x, plhs[1]->d.data(), sizeof(double) * 2 * n * B);
    std::memcpy(cost, plhs[2]->d.data(), sizeof(double) * max_outer * B);
    std::memcpy(eu, plhs[3]->d.data(), sizeof(double) * max_outer * B);
    std::memcpy(iters, plhs[4]->i32.data(), sizeof(int) * B);
    std::memcpy(status, plhs[5]->i32.data(), sizeof(int) * B);
    *qp_steps = plhs[6]->d[0];
  } catch (const mex_error &e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", e.id.c_str(), e.msg.c_str());
    rc = 1;
  }
  for (mxArray *a : g_all) delete a;
  g_all.clear();
  return rc;
}
