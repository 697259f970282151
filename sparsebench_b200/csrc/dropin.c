/* Link-time drop-in shims: the reference links exactly one matrix-<FMT>.o that defines convertMatrix and
 * spMVM for the -D<FMT> Matrix typedef (Makefile:20,32-34; matrix.h:14-22), and solver.h:11 declares
 * solveCG on that same typedef. libsparsebench_b200_<FMT>.so, built from this file with -DCRS, -DSCS or
 * -DCCRS, exports those three bare names and forwards to the format-prefixed entry points of
 * libsparsebench_b200.so; every other reference symbol on the path (waxpby, ddot, commPartition,
 * commExchange, commReduction, commInit, commFinalize, allocate, getTimeStamp, matrixGenerate) is
 * exported by the core library under its own name already. */
#include "sparsebench_b200.h"

#undef convertMatrix
#undef spMVM
#undef solveCG

#if defined(CRS)
#define SB_(name) sbCRS_##name
#elif defined(SCS)
#define SB_(name) sbSCS_##name
#elif defined(CCRS)
#define SB_(name) sbCCRS_##name
#else
#error "define CRS, SCS or CCRS"
#endif

void convertMatrix(Matrix* m, GMatrix* im) { SB_(convertMatrix)(m, im); }          /* matrix.h:57 */
void spMVM(Matrix* m, const CG_FLOAT* x, CG_FLOAT* y) { SB_(spMVM)(m, x, y); }     /* solver.h:13 */
int solveCG(Comm* comm, Parameter* param, Matrix* m) { return SB_(solveCG)(comm, param, m); }   /* solver.h:11 */
