// matio.cpp -- data formats on the input side of the path: PETSc binary matrices / vectors
// (MatLoad / VecLoad through a binary viewer, the format PETSc 3.7.6's MatView writes) and
// MatrixMarket coordinate files.  SURVEY 8(f) rank 4: lets external matrices reach the kernels
// through the same MatSetValues / MatAssemblyEnd_SeqAIJ route as the reference problem.
//
// PETSc binary [P376]: big-endian; Mat = int32 classid 1211216, M, N, nz, int32 rowlen[M],
// int32 col[nz], float64 val[nz]; Vec = int32 classid 1211214, n, float64 val[n].
#include <cstdlib>
#include <string>
#include <vector>

#include "b200_aij.h"

struct _p_PetscViewer { FILE *f; std::string name; };

namespace {
uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
bool read_i32(FILE *f, PetscInt *out, size_t n)
{
  std::vector<unsigned char> buf(4 * n);
  if (n && fread(buf.data(), 4, n, f) != n) return false;
  for (size_t i = 0; i < n; ++i) out[i] = (PetscInt)be32(&buf[4 * i]);
  return true;
}
bool read_f64(FILE *f, double *out, size_t n)
{
  std::vector<unsigned char> buf(8 * n);
  if (n && fread(buf.data(), 8, n, f) != n) return false;
  for (size_t i = 0; i < n; ++i) {
    uint64_t v = ((uint64_t)be32(&buf[8 * i]) << 32) | be32(&buf[8 * i + 4]);
    memcpy(&out[i], &v, 8);
  }
  return true;
}
constexpr PetscInt MAT_FILE_CLASSID = 1211216, VEC_FILE_CLASSID = 1211214;
}  // namespace

extern "C" PetscErrorCode PetscViewerBinaryOpen(MPI_Comm comm, const char name[], PetscFileMode mode, PetscViewer *v)
{
  if (mode != FILE_MODE_READ) SETERRQ(comm, PETSC_ERR_SUP, "binary viewer: only FILE_MODE_READ");
  FILE *f = fopen(name, "rb");
  if (!f) SETERRQ1(comm, PETSC_ERR_FILE_OPEN, "Cannot open file %s", name);
  *v = new _p_PetscViewer{f, name};
  return 0;
}
extern "C" PetscErrorCode PetscViewerDestroy(PetscViewer *v)
{
  if (v && *v) { if ((*v)->f) fclose((*v)->f); delete *v; *v = NULL; }
  return 0;
}

// MatLoad_SeqAIJ [P376]: header, row lengths (exact preallocation), then row by row MatSetValues
extern "C" PetscErrorCode MatLoad(Mat *newmat, PetscViewer v)
{
  PetscInt hdr[4];
  PetscErrorCode ierr;
  if (!read_i32(v->f, hdr, 4)) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_FILE_UNEXPECTED, "%s: truncated header", v->name.c_str());
  if (hdr[0] != MAT_FILE_CLASSID) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_FILE_UNEXPECTED, "%s: not a PETSc matrix file", v->name.c_str());
  const PetscInt M = hdr[1], N = hdr[2], nz = hdr[3];
  if (M < 0 || N < 0 || nz < 0) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_FILE_UNEXPECTED, "negative sizes in matrix header");
  std::vector<PetscInt> len((size_t)M), cols((size_t)nz);
  std::vector<double>   vals((size_t)nz);
  if (!read_i32(v->f, len.data(), M) || !read_i32(v->f, cols.data(), nz) || !read_f64(v->f, vals.data(), nz))
    SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_FILE_UNEXPECTED, "%s: truncated matrix data", v->name.c_str());
  long long tot = 0;
  for (PetscInt i = 0; i < M; ++i) { if (len[i] < 0) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_FILE_UNEXPECTED, "negative row length"); tot += len[i]; }
  if (tot != nz) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_FILE_UNEXPECTED, "row lengths sum to %d, header says %d", (int)tot, nz);
  ierr = MatCreateSeqAIJ(PETSC_COMM_SELF, M, N, 0, len.data(), newmat);CHKERRQ(ierr);
  PetscInt at = 0;
  for (PetscInt i = 0; i < M; ++i) {
    ierr = MatSetValues(*newmat, 1, &i, len[i], cols.data() + at, vals.data() + at, INSERT_VALUES);CHKERRQ(ierr);
    at += len[i];
  }
  ierr = MatAssemblyBegin(*newmat, MAT_FINAL_ASSEMBLY);CHKERRQ(ierr);
  ierr = MatAssemblyEnd(*newmat, MAT_FINAL_ASSEMBLY);CHKERRQ(ierr);
  return 0;
}

extern "C" PetscErrorCode VecLoad(Vec *newvec, PetscViewer v)
{
  PetscInt hdr[2];
  PetscErrorCode ierr;
  if (!read_i32(v->f, hdr, 2) || hdr[0] != VEC_FILE_CLASSID || hdr[1] < 0)
    SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_FILE_UNEXPECTED, "%s: not a PETSc vector file", v->name.c_str());
  ierr = VecCreateSeq(PETSC_COMM_SELF, hdr[1], newvec);CHKERRQ(ierr);
  PetscScalar *a;
  ierr = VecGetArray(*newvec, &a);CHKERRQ(ierr);
  const bool ok = read_f64(v->f, a, hdr[1]);
  ierr = VecRestoreArray(*newvec, &a);CHKERRQ(ierr);
  if (!ok) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_FILE_UNEXPECTED, "%s: truncated vector data", v->name.c_str());
  return 0;
}

// MatrixMarket "matrix coordinate {real|integer|pattern} {general|symmetric|skew-symmetric}";
// duplicates are summed (ADD_VALUES), the symmetric half is mirrored.
extern "C" PetscErrorCode MatLoadMatrixMarketB200(const char path[], Mat *newmat)
{
  FILE *f = fopen(path, "r");
  if (!f) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_FILE_OPEN, "Cannot open file %s", path);
  char line[1024], obj[64], fmt[64], field[64], sym[64];
  PetscErrorCode ierr = 0;
  auto fail = [&](const char *msg) { fclose(f); return PetscError(PETSC_COMM_SELF, __LINE__, "MatLoadMatrixMarketB200", __FILE__, PETSC_ERR_FILE_UNEXPECTED, "%s: %s", path, msg); };
  if (!fgets(line, sizeof line, f) || sscanf(line, "%%%%MatrixMarket %63s %63s %63s %63s", obj, fmt, field, sym) != 4) return fail("missing %%MatrixMarket banner");
  for (char *p : {obj, fmt, field, sym}) for (; *p; ++p) *p = (char)tolower(*p);
  if (strcmp(obj, "matrix") || strcmp(fmt, "coordinate")) return fail("only 'matrix coordinate' files");
  const bool pattern = !strcmp(field, "pattern");
  if (!pattern && strcmp(field, "real") && strcmp(field, "integer") && strcmp(field, "double")) return fail("only real / integer / pattern fields");
  const bool symm = !strcmp(sym, "symmetric"), skew = !strcmp(sym, "skew-symmetric");
  if (!symm && !skew && strcmp(sym, "general")) return fail("only general / symmetric / skew-symmetric");
  do { if (!fgets(line, sizeof line, f)) return fail("missing size line"); } while (line[0] == '%' || line[0] == '\n');
  long long M, N, nz;
  if (sscanf(line, "%lld %lld %lld", &M, &N, &nz) != 3 || M < 0 || N < 0 || nz < 0 || M > 2147483647LL || N > 2147483647LL) return fail("bad size line");
  std::vector<PetscInt> ri((size_t)nz), ci((size_t)nz);
  std::vector<double>   va((size_t)nz, 1.0);
  std::vector<PetscInt> cnt((size_t)M, 0);
  for (long long k = 0; k < nz; ++k) {
    long long i, j;
    double    v = 1.0;
    if (!fgets(line, sizeof line, f)) return fail("fewer entries than announced");
    const int got = pattern ? sscanf(line, "%lld %lld", &i, &j) : sscanf(line, "%lld %lld %lf", &i, &j, &v);
    if (got != (pattern ? 2 : 3) || i < 1 || j < 1 || i > M || j > N) return fail("bad entry");
    ri[k] = (PetscInt)(i - 1); ci[k] = (PetscInt)(j - 1); va[k] = v;
    cnt[i - 1]++;
    if ((symm || skew) && i != j) cnt[j - 1]++;
  }
  fclose(f);
  ierr = MatCreateSeqAIJ(PETSC_COMM_SELF, (PetscInt)M, (PetscInt)N, 0, cnt.data(), newmat);CHKERRQ(ierr);
  for (long long k = 0; k < nz; ++k) {
    ierr = MatSetValues(*newmat, 1, &ri[k], 1, &ci[k], &va[k], ADD_VALUES);CHKERRQ(ierr);
    if ((symm || skew) && ri[k] != ci[k]) {
      const double w = skew ? -va[k] : va[k];
      ierr = MatSetValues(*newmat, 1, &ci[k], 1, &ri[k], &w, ADD_VALUES);CHKERRQ(ierr);
    }
  }
  ierr = MatAssemblyBegin(*newmat, MAT_FINAL_ASSEMBLY);CHKERRQ(ierr);
  ierr = MatAssemblyEnd(*newmat, MAT_FINAL_ASSEMBLY);CHKERRQ(ierr);
  return 0;
}
