/* Minimal stand-in for MATLAB's mex.h: ONLY for compile-checking matlab/gnssacq_mex.c in an image
 * without MATLAB/Octave (tests/test_cabi.py).  Declarations only; nothing links against it. */
#ifndef STUB_MEX_H_
#define STUB_MEX_H_
#include <stddef.h>
#include <stdio.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
int mxIsChar(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
int mxIsStruct(const mxArray*); int mxIsInt8(const mxArray*); int mxIsInt16(const mxArray*); int mxIsDouble(const mxArray*);
size_t mxGetM(const mxArray*); size_t mxGetN(const mxArray*);
mxArray* mxGetField(const mxArray*, mwSize, const char*);
size_t mxGetNumberOfElements(const mxArray*); size_t mxGetElementSize(const mxArray*);
double mxGetScalar(const mxArray*); double* mxGetPr(const mxArray*); void* mxGetData(const mxArray*);
void* mxMalloc(size_t); void mxFree(void*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexLock(void); int mexAtExit(void (*)(void));
#endif
