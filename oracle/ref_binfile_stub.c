/* TEST INFRASTRUCTURE ONLY -- stands in for the reference's src/matrixBinfile.c, which includes mpi.h
 * unconditionally (matrixBinfile.c:8) and therefore cannot be compiled in a non-MPI build. Only main.c:51,76
 * reference these two entry points (.bmx files), which no test uses. */
#include <stdio.h>
#include <stdlib.h>

#include "matrixBinfile.h"

void matrixBinWrite(GMatrix* m, Comm* c, char* filename) { (void)m; (void)c; (void)filename; fprintf(stderr, ".bmx not available in this build\n"); exit(EXIT_FAILURE); }
void matrixBinRead(GMatrix* m, Comm* c, char* filename) { (void)m; (void)c; (void)filename; fprintf(stderr, ".bmx not available in this build\n"); exit(EXIT_FAILURE); }
