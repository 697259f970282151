// MatrixMarket input path (replaces MMMatrixRead / matrixConvertfromMM, matrix.c:123-269).
// Host-side I/O, run once: the arrays it produces are HOST arrays (posix_memalign, like the reference's
// allocate.c) -- convertMatrix and commPartition accept host GMatrix input and move it to the device.
//
// Behaviour follows the reference: coordinate format, real / integer / pattern values, general or symmetric
// storage (off-diagonal entries are mirrored right behind the entry they come from, matrix.c:208-212), then a
// sort by column followed by a STABLE sort by row (matrix.c:220-228), i.e. rows ascending, columns ascending
// inside a row, duplicates in file order. Anything else is rejected with the reference's messages + exit.
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "sparsebench_b200.h"

namespace {

void* hostAlloc(size_t bytes)
{
  void* p = nullptr;
  if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) {   // allocate.c:12-36 (ARRAY_ALIGNMENT = 64)
    fprintf(stderr, "Error: Insufficient memory to fulfill the request\n");
    exit(EXIT_FAILURE);
  }
  return p;
}

std::string lower(const char* s)
{
  std::string r(s);
  for (char& c : r) c = (char)tolower((unsigned char)c);
  return r;
}

} // namespace

extern "C" {

void MMMatrixRead(MMMatrix* m, char* filename)
{
  FILE* f = fopen(filename, "r");
  if (!f) {
    printf("Unable to open file.\n");                                  // matrix.c:129-132
    exit(EXIT_FAILURE);
  }
  // banner: %%MatrixMarket matrix coordinate <field> <symmetry>
  char line[1100], tag[64], object[64], format[64], field[64], symmetry[64];
  if (!fgets(line, sizeof(line), f) ||
      sscanf(line, "%63s %63s %63s %63s %63s", tag, object, format, field, symmetry) != 5 ||
      lower(tag) != "%%matrixmarket") {
    printf("Could not process Matrix Market banner.\n");               // :134-137
    exit(EXIT_FAILURE);
  }
  const std::string obj = lower(object), fmt = lower(format), fld = lower(field), sym = lower(symmetry);
  const bool isMatrix = obj == "matrix", isSparse = fmt == "coordinate";
  const bool isReal = fld == "real", isInteger = fld == "integer", isPattern = fld == "pattern";
  const bool isSymmetric = sym == "symmetric", isGeneral = sym == "general";
  if (!((isReal || isPattern || isInteger) && isMatrix && isSparse)) {
    fprintf(stderr, "Sorry, this application does not support ");     // :139-145
    fprintf(stderr, "Market Market type: [%s %s %s %s]\n", obj.c_str(), fmt.c_str(), fld.c_str(), sym.c_str());
    exit(EXIT_FAILURE);
  }
  if (!(isSymmetric || isGeneral)) {
    printf("The matrix market file provided is not supported.\n Reason :\n");   // :156-171
    printf(" * matrix has to be symmetric\n");
    exit(EXIT_FAILURE);
  }
  // size line: first line that is not a comment
  int M = 0, N = 0, nz = 0;
  do {
    if (!fgets(line, sizeof(line), f)) exit(EXIT_FAILURE);
  } while (line[0] == '%');
  while (sscanf(line, "%d %d %d", &M, &N, &nz) != 3) {                 // blank lines before the size line
    if (!fgets(line, sizeof(line), f)) exit(EXIT_FAILURE);
  }
  printf("Read matrix %s with %d non zeroes and %d rows\n", filename, nz, M);   // :178

  std::vector<MMEntry> e;
  e.reserve((size_t)nz * (isSymmetric ? 2 : 1));
  for (int i = 0; i < nz; i++) {
    int row = 0, col = 0;
    double v = 1.0;
    int got;
    if (isPattern) got = fscanf(f, "%d %d\n", &row, &col) + 1;
    else got = fscanf(f, "%d %d %lg\n", &row, &col, &v);
    if (got != 3) {
      fprintf(stderr, "MMMatrixRead: malformed entry %d in %s\n", i + 1, filename);
      exit(EXIT_FAILURE);
    }
    row--;                                                             // :201-202
    col--;
    e.push_back(MMEntry { row, col, v });
    if (isSymmetric && row != col) e.push_back(MMEntry { col, row, v });   // :208-212
  }
  fclose(f);
  std::stable_sort(e.begin(), e.end(), [](const MMEntry& a, const MMEntry& b) { return a.col < b.col; });   // :220
  std::stable_sort(e.begin(), e.end(), [](const MMEntry& a, const MMEntry& b) { return a.row < b.row; });   // :224
  m->entries = (MMEntry*)hostAlloc(sizeof(MMEntry) * e.size());
  if (!e.empty()) memcpy(m->entries, e.data(), sizeof(MMEntry) * e.size());
  m->nr = M;                                                           // :215-217
  m->nnz = (int)e.size();
  m->count = e.size();
}

void matrixConvertfromMM(MMMatrix* mm, GMatrix* m)
{
  m->startRow = (CG_UINT)mm->startRow;                                 // matrix.c:233-239
  m->stopRow = (CG_UINT)mm->stopRow;
  m->totalNr = (CG_UINT)mm->totalNr;
  m->totalNnz = (CG_UINT)mm->totalNnz;
  m->nr = (CG_UINT)mm->nr;
  m->nc = (CG_UINT)mm->nr;
  m->nnz = (CG_UINT)mm->nnz;
  m->entries = (Entry*)hostAlloc(sizeof(Entry) * (size_t)m->nnz);
  m->rowPtr = (CG_UINT*)hostAlloc(sizeof(CG_UINT) * ((size_t)m->nr + 1));
  memset(m->entries, 0, sizeof(Entry) * (size_t)m->nnz);              // defined padding bytes
  std::vector<CG_UINT> perRow((size_t)m->nr, 0);
  for (size_t i = 0; i < mm->count; i++) perRow[(size_t)(mm->entries[i].row - mm->startRow)]++;   // :253-255
  m->rowPtr[0] = 0;
  for (CG_UINT r = 0; r < m->nr; r++) m->rowPtr[r + 1] = m->rowPtr[r] + perRow[r];                // :259-261
  for (size_t i = 0; i < mm->count; i++) {                             // entries are already in row order (:263-266)
    m->entries[i].val = (CG_FLOAT)mm->entries[i].val;
    m->entries[i].col = (CG_UINT)mm->entries[i].col;
  }
}

} // extern "C"

// ------------------------------------------------------------------------------------------- .bmx binary files
// Replaces matrixBinWrite / matrixBinRead (matrixBinfile.c:38-236), which go through MPI-IO. File layout, as that
// code writes it: 24 bytes "# SparseBench DataFile" (NUL padded, :56-61), u32 totalNr, u32 totalNnz, u32
// rowPtr[totalNr+1] (global entry offsets, :75-82), then totalNnz records {u32 col; f32 val} (:19-35, :92-103).
// Values are float32 on disk. Every rank reads its own row block with plain POSIX I/O: rows split like sizeOfRank
// (:14-17, :157-164), row pointers made local by subtracting the block's first entry offset (:189-205). Host arrays.
// Parity: pinned by the reference itself -- its matrixBinfile.c compiled against the test-only MPI shim (whose MPI-IO
// subset runs on POSIX files) wrote tests/golden/reference_fixtures/klein_ref.bmx and read it back on 1, 2, 3 and 7
// ranks (tests/golden/make_file_golden.py); the file written here is byte-identical, the row blocks read here equal.
namespace {

constexpr size_t kBmxHeader = 24;
struct FEntry {               // matrixBinfile.h:11-14
  unsigned int col;
  float val;
};

void readExact(FILE* f, void* dst, size_t bytes, const char* what)
{
  if (fread(dst, 1, bytes, f) != bytes) {
    printf("ERROR reading %s!\n", what);           // matrixBinfile.c:125-127,141-148
    exit(EXIT_FAILURE);
  }
}

} // namespace

extern "C" {

void matrixBinWrite(GMatrix* m, Comm* c, char* filename)
{
  if (c->size > 1) {
    fprintf(stderr, "ERROR: Matrix writing only supported for single rank\n");   // matrixBinfile.c:42-45
    return;
  }
  FILE* f = fopen(filename, "wb");
  if (!f) {
    fprintf(stderr, "ERROR: cannot open %s for writing\n", filename);
    exit(EXIT_FAILURE);
  }
  printf("Writing matrix to %s\n", filename);
  char header[kBmxHeader];
  memset(header, 0, sizeof(header));
  strncpy(header, "# SparseBench DataFile", sizeof(header) - 1);
  const unsigned int totalNr = m->totalNr, totalNnz = m->rowPtr[m->nr];   // entries actually stored
  bool ok = fwrite(header, 1, kBmxHeader, f) == kBmxHeader && fwrite(&totalNr, 4, 1, f) == 1 && fwrite(&totalNnz, 4, 1, f) == 1 &&
            fwrite(m->rowPtr, 4, (size_t)totalNr + 1, f) == (size_t)totalNr + 1;
  std::vector<FEntry> out((size_t)totalNnz);
  for (size_t i = 0; i < out.size(); i++) {
    out[i].col = (unsigned int)m->entries[i].col;
    out[i].val = (float)m->entries[i].val;            // :99-102
  }
  ok = ok && fwrite(out.data(), sizeof(FEntry), out.size(), f) == out.size();
  if (fclose(f) != 0 || !ok) {
    fprintf(stderr, "ERROR: writing %s failed\n", filename);
    exit(EXIT_FAILURE);
  }
}

void matrixBinRead(GMatrix* m, Comm* c, char* filename)
{
  FILE* f = fopen(filename, "rb");
  if (!f) {
    fprintf(stderr, "ERROR: cannot open %s\n", filename);
    exit(EXIT_FAILURE);
  }
  if (c->rank == 0) printf("Reading matrix from %s\n", filename);
  char header[kBmxHeader];
  unsigned int totalNr = 0, totalNnz = 0;
  readExact(f, header, kBmxHeader, "header");
  readExact(f, &totalNr, 4, "unsigned");
  readExact(f, &totalNnz, 4, "unsigned");
  m->totalNr = totalNr;
  m->totalNnz = totalNnz;
  // row block of this rank (matrixBinfile.c:157-176)
  unsigned int startRow = 0, numRows = 0;
  for (int i = 0; i <= c->rank; i++) {
    startRow += numRows;
    numRows = totalNr / (unsigned int)c->size + ((totalNr % (unsigned int)c->size > (unsigned int)i) ? 1u : 0u);
  }
  m->nr = numRows;
  m->nc = numRows;
  m->startRow = startRow;
  m->stopRow = startRow + numRows - 1;
  m->rowPtr = (CG_UINT*)hostAlloc(sizeof(CG_UINT) * ((size_t)numRows + 1));
  const long rowPtrAt = (long)(kBmxHeader + 8);
  if (fseek(f, rowPtrAt + 4L * (long)startRow, SEEK_SET) != 0) exit(EXIT_FAILURE);
  readExact(f, m->rowPtr, 4 * ((size_t)numRows + 1), "rowptr");
  const unsigned int entryOffset = m->rowPtr[0];        // = non-zeros of all lower ranks (:195-205)
  for (unsigned int i = 0; i <= numRows; i++) m->rowPtr[i] -= entryOffset;
  m->nnz = m->rowPtr[numRows];
  std::vector<FEntry> in((size_t)m->nnz);
  const long entriesAt = rowPtrAt + 4L * ((long)totalNr + 1);
  if (fseek(f, entriesAt + 8L * (long)entryOffset, SEEK_SET) != 0) exit(EXIT_FAILURE);
  readExact(f, in.data(), sizeof(FEntry) * in.size(), "entries");
  fclose(f);
  m->entries = (Entry*)hostAlloc(sizeof(Entry) * (size_t)(m->nnz ? m->nnz : 1));
  memset(m->entries, 0, sizeof(Entry) * (size_t)(m->nnz ? m->nnz : 1));
  for (size_t i = 0; i < in.size(); i++) {              // :229-232
    m->entries[i].col = (CG_UINT)in[i].col;
    m->entries[i].val = (CG_FLOAT)in[i].val;
  }
}

} // extern "C"
