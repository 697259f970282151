/* The two reference entry points main.c references that this package does not provide (binary .bmx matrix files:
 * MPI-IO, float32 values; out of scope). */
#include <stdio.h>
#include <stdlib.h>

#include "comm.h"

void matrixBinWrite(GMatrix* m, Comm* c, char* filename);
void matrixBinRead(GMatrix* m, Comm* c, char* filename);

void matrixBinWrite(GMatrix* m, Comm* c, char* filename)
{
  (void)m; (void)c;
  fprintf(stderr, "sparsebench_b200: writing %s: the .bmx format is not supported by this build\n", filename);
  exit(EXIT_FAILURE);
}

void matrixBinRead(GMatrix* m, Comm* c, char* filename)
{
  (void)m; (void)c;
  fprintf(stderr, "sparsebench_b200: reading %s: the .bmx format is not supported by this build\n", filename);
  exit(EXIT_FAILURE);
}
