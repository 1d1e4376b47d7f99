// mex.h -- minimal stand-in for MATLAB's MEX API (test infrastructure only).
// MATLAB is not installed in the build container, so matlab/cfs_mex.cpp cannot be built with `mex`.  This stub implements
// exactly the subset of the mx*/mex* API the gateway uses (column-major double / int32 matrices, structs, cells, char rows)
// so that the gateway is compiled and RUN by the test-suite (tests/test_mex_gateway.py through tests/mex_stub/harness.cpp).
#pragma once
#include <cstddef>
#include <map>
#include <string>
#include <vector>

typedef enum { mxREAL = 0 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxSTRUCT_CLASS = 2, mxCELL_CLASS = 1, mxCHAR_CLASS = 4 } mxClassID;

struct mxArray {
  mxClassID cls = mxDOUBLE_CLASS;
  size_t m = 0, n = 0;
  std::vector<double> d;                       // mxDOUBLE_CLASS
  std::vector<int> i32;                        // mxINT32_CLASS
  std::map<std::string, mxArray *> fields;     // mxSTRUCT_CLASS (1 x 1)
  std::vector<mxArray *> cells;                // mxCELL_CLASS
  std::string str;                             // mxCHAR_CLASS
};

double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
double mxGetScalar(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
bool mxIsEmpty(const mxArray *a);
bool mxIsDouble(const mxArray *a);
bool mxIsStruct(const mxArray *a);
bool mxIsCell(const mxArray *a);
bool mxIsChar(const mxArray *a);
mxArray *mxGetField(const mxArray *s, size_t index, const char *name);
mxArray *mxGetCell(const mxArray *c, size_t index);
int mxGetString(const mxArray *a, char *buf, size_t buflen);
mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateNumericMatrix(size_t m, size_t n, mxClassID cls, mxComplexity c);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);  // throws mex_error
int mexAtExit(void (*fn)(void));

struct mex_error {
  std::string id, msg;
};

extern "C" void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
