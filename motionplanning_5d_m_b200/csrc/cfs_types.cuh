// cfs_types.cuh -- shared POD types of libcfs_b200 (host + device).
#pragma once
#include <cstdint>

#define CFS_MAXL 6      // links per robot table (robotproperty2.m: nlink = 6; 5 used)
#define CFS_MAX_OBS 32  // obstacle capsules staged per table copy

// One DH link, pre-digested on the host so that the device evaluates exactly the products of
// Lib/functions/CapPos.m:13-16 (R, T from theta,d,a,alpha) with the alpha trigs as constants.
// 2L robot (Lib/2L/CapPos2.m:16-29): ca=1, sa=0, a=0 and (tx,ty,dz) = robot.T(:,i).
struct LinkTab {
  double ca, sa;      // cos(alpha), sin(alpha)            (literal DH constants 1.5708 / 3.1416 kept)
  double a;           // DH a
  double tx, ty, dz;  // constant translation: (0,0,d) for DH robots, robot.T(:,i) for 2L
  double th_off;      // joint offset: -pi/2 on M200i joint 2 (dist_arm_3D_200i_2.m:11)
  double pad_;
  double cap[2][3];   // cap{i}.p(:,k)
};  // 14 doubles = 112 B

struct ObsTab {
  double s[3];    // obs.l(:,1)                                                    | box: min corner
  double d2[3];   // obs.l(:,2) - obs.l(:,1)   (distLinSeg.m:26, computed once)    | box: max corner
  double D2;      // sum(d2.^2)                (distLinSeg.m:30)
  double D, eps;  // obs.D, obs.epsilon
  double rD2;     // RN(1/D2) (0 when D2 == 0): x / D2 is evaluated as q0 = x*rD2, q = q0 + (x - q0*D2)*rD2 with FMAs,
                  // which returns the correctly rounded quotient (Markstein) for 3 instructions instead of ~18
  int kind;       // CFS_OBS_CAPSULE: capsule axis (every obstacle of the reference); CFS_OBS_BOX: solid axis-aligned box
  int pad_[3];    //   (N3 extension, obs{j}.shape = 'box': e.g. the bounding box of an STL part of map/)
};  // 12 doubles = 96 B
#define CFS_OBS_CAPSULE 0
#define CFS_OBS_BOX 1

// Staged into shared memory with one TMA bulk copy (cp.async.bulk) per CTA.  sizeof % 16 == 0.
struct DevTables {
  LinkTab link[CFS_MAXL];  // 672 B
  double base[3];
  double dt;
  int kind, nj, nobs, pad_;
  ObsTab obs[CFS_MAX_OBS];
};
static_assert(sizeof(LinkTab) % 16 == 0, "LinkTab must be 16B granular");
static_assert(sizeof(ObsTab) % 16 == 0, "ObsTab must be 16B granular");
static_assert(sizeof(DevTables) % 16 == 0, "DevTables must be 16B granular");
#define CFS_TAB_HEADER_BYTES (sizeof(LinkTab) * CFS_MAXL + 4 * sizeof(double) + 4 * sizeof(int))

// DERIVEST constants (derivest.m defaults), computed once on the host in FP64.
#define DV_NDEL 26
#define DV_NE 23
#define DV_NEST 19
struct DerivestTab {
  double delta[DV_NDEL];  // 100 * sr.^(0:-1:-25)        derivest.m:238
  double fdarule[2];      // [1 0]/fdamat(sr,1,2)        derivest.m:282
  double rmat[4][3];      // rombextrap rmat             derivest.m:493-499
  double q[4][3];         // economy QR of rmat          derivest.m:510
  double rr[3][3];
  double errfac;          // 12.7062047361747*sqrt(cov1(1))  derivest.m:523-525
};
static_assert(sizeof(DerivestTab) % 16 == 0, "DerivestTab must be 16B granular");

#define CFS_NUMJAC_EPS 1e-5  // Lib/functions/num_jac.m:6
// cos / sin of the half step eps/2 = 0x1.4f8b588e368f1p-18 (correctly rounded): the perturbed joints' sin/cos follow from
// the unperturbed pair by the angle-addition formulas (2 FMAs + 2 multiplies each, error <= 1 ulp -- the size of the
// argument rounding in x(i)+eps/2 itself, num_jac.m:11) instead of two more sincos evaluations per joint
#define CFS_NUMJAC_COSH 0x1.ffffffffe4832p-1
#define CFS_NUMJAC_SINH 0x1.4f8b588e308ddp-18
#define CFS_TOUCH_TOL 0.0001 // dist_arm_3D_Heu_2.m:22
