// minipetsc.cpp -- sys / Vec / Mat(SeqAIJ) of the PETSc 3.7.6 API slice (include/b200_petsc.h).
//
// Mirrors what the reference application calls (src/main_ksp.cpp, src/helper.cpp) around the hot
// path.  Host-side integer work (MatSetValues' sorted insertion, preallocation bookkeeping, the
// compaction in MatAssemblyEnd_SeqAIJ, MatZeroRowsColumns) follows PETSc's published algorithms
// [P376] so that the assembled CSR is bit-identical to what the oracle's restatement of
// src/helper.cpp produces.  Vector arithmetic runs on the GPU through the C ABI (no CPU fallback);
// only set/copy/array access and the sequential VecSum (setup code whose rounding must match the
// host loop) touch host memory.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "b200_aij.h"
// (after b200_aij.h: the symbols header only forward-declares Mat/Vec when PETSc types are absent)
#include "../../../include/b200_petsc_symbols.h"
#include "../../../include/b200_seqaij.h"

// ---------------------------------------------------------------------------------------------
// errors, MPI stubs, options, time, print
// ---------------------------------------------------------------------------------------------
extern "C" PetscErrorCode PetscError(MPI_Comm, int line, const char *func, const char *file, PetscErrorCode code, const char *fmt, ...)
{
  char    buf[1024] = "";
  va_list ap;
  va_start(ap, fmt);
  if (fmt) vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  const char *lib = b200_last_error();
  fprintf(stderr, "[0]PETSC ERROR: %s() line %d in %s: error %d %s%s%s\n", func, line, file, code, buf,
          (lib && *lib) ? " | b200: " : "", (lib && *lib) ? lib : "");
  return code ? code : PETSC_ERR_LIB;
}

extern "C" int MPI_Init(int *, char ***) { return 0; }
extern "C" int MPI_Finalize(void) { return 0; }
extern "C" int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
extern "C" int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }

namespace {
std::map<std::string, std::string> g_options;
double                             g_flops = 0.0;
bool                               g_have_device = false, g_device_probed = false;

bool have_device()
{
  if (!g_device_probed) {
    g_device_probed = true;
    g_have_device   = (b200_device_sm_count() > 0);
  }
  return g_have_device;
}

void add_option(const std::string &k, const std::string &v) { g_options[k[0] == '-' ? k.substr(1) : k] = v; }

const std::string *find_option(const char *pre, const char *name)
{
  std::string key = name ? name : "";
  if (!key.empty() && key[0] == '-') key = key.substr(1);
  if (pre) key = std::string(pre) + key;
  auto it = g_options.find(key);
  return it == g_options.end() ? nullptr : &it->second;
}
}  // namespace

extern "C" PetscErrorCode PetscInitialize(int *argc, char ***args, const char file[], const char[])
{
  if (argc && args) {
    for (int i = 1; i < *argc; ++i) {
      const char *a = (*args)[i];
      if (a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
        std::string v;
        if (i + 1 < *argc) {
          const char *b = (*args)[i + 1];
          if (!(b[0] == '-' && !((b[1] >= '0' && b[1] <= '9') || b[1] == '.'))) { v = b; ++i; }
        }
        add_option(a, v);
      }
    }
  }
  if (file) return PetscOptionsInsertFile(PETSC_COMM_WORLD, NULL, file, PETSC_FALSE);
  return 0;
}
extern "C" PetscErrorCode PetscFinalize(void) { g_options.clear(); return 0; }
extern "C" PetscErrorCode PetscOptionsClear(PetscOptions) { g_options.clear(); return 0; }
extern "C" PetscErrorCode PetscOptionsSetValue(PetscOptions, const char name[], const char value[])
{
  add_option(name, value ? value : "");
  return 0;
}
extern "C" PetscErrorCode PetscOptionsGetString(PetscOptions, const char pre[], const char name[], char str[], size_t len, PetscBool *set)
{
  const std::string *v = find_option(pre, name);
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  if (v && str && len) { strncpy(str, v->c_str(), len - 1); str[len - 1] = 0; }
  else if (str && len) str[0] = 0;
  return 0;
}
extern "C" PetscErrorCode PetscOptionsGetInt(PetscOptions, const char pre[], const char name[], PetscInt *val, PetscBool *set)
{
  const std::string *v = find_option(pre, name);
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  if (v && val) *val = atoi(v->c_str());
  return 0;
}
extern "C" PetscErrorCode PetscOptionsGetReal(PetscOptions, const char pre[], const char name[], PetscReal *val, PetscBool *set)
{
  const std::string *v = find_option(pre, name);
  if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
  if (v && val) *val = atof(v->c_str());
  return 0;
}
// options file: one "-name value" per line, '#' comments (configs/PETSc_SolverOptions_GAMG.info)
extern "C" PetscErrorCode PetscOptionsInsertFile(MPI_Comm comm, PetscOptions, const char file[], PetscBool require)
{
  if (!file || !*file) return 0;
  FILE *f = fopen(file, "r");
  if (!f) {
    if (require) SETERRQ1(comm, PETSC_ERR_FILE_OPEN, "Unable to open options file %s", file);
    return 0;
  }
  char line[4096];
  while (fgets(line, sizeof line, f)) {
    char *h = strchr(line, '#');
    if (h) *h = 0;
    char name[2048] = "", val[2048] = "";
    int  k = sscanf(line, " %2047s %2047s", name, val);
    if (k >= 1 && name[0] == '-') add_option(name, k == 2 ? val : "");
  }
  fclose(f);
  return 0;
}
extern "C" PetscErrorCode PetscTime(PetscLogDouble *t)
{
  *t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return 0;
}
extern "C" PetscErrorCode PetscPrintf(MPI_Comm, const char fmt[], ...)
{
  va_list ap;
  va_start(ap, fmt);
  vprintf(fmt, ap);
  va_end(ap);
  fflush(stdout);
  return 0;
}
extern "C" PetscErrorCode PetscLogFlops(PetscLogDouble f) { g_flops += f; return 0; }
extern "C" PetscErrorCode PetscGetFlops(PetscLogDouble *f) { *f = g_flops; return 0; }

// ---------------------------------------------------------------------------------------------
// Vec: page-locked host array + lazily allocated device mirror
// ---------------------------------------------------------------------------------------------
#define CUDA_CHK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_LIB, "CUDA: %s", cudaGetErrorString(_e)); } while (0)
#define B200_CHK(expr) do { int _r = (expr); if (_r) return PetscError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, _r, " "); } while (0)

namespace {
struct VecPriv { bool pinned; };
std::map<Vec, VecPriv> g_vecpriv;

// The host array appears on the first host access (zero-filled, page-locked when a device is
// present): the work vectors of a solve live in HBM only and never pay for pinned memory.
// Until then array == NULL with host_valid set means "all zeros".
PetscErrorCode vec_host_alloc(Vec x)
{
  if (x->array) return 0;
  const size_t bytes = sizeof(PetscScalar) * (size_t)PetscMax(x->n, 1);
  void *p = NULL;
  bool  pinned = false;
  if (have_device() && b200_host_alloc(&p, bytes) == 0) pinned = true;
  else p = malloc(bytes);
  if (!p) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_MEM, "out of memory");
  memset(p, 0, bytes);
  x->array = (PetscScalar *)p;
  g_vecpriv[x] = VecPriv{pinned};
  return 0;
}
PetscErrorCode vec_to_host(Vec x)
{
  PetscErrorCode ierr = vec_host_alloc(x);CHKERRQ(ierr);
  if (!x->host_valid) {
    CUDA_CHK(cudaMemcpy(x->array, x->d_array, sizeof(PetscScalar) * (size_t)x->n, cudaMemcpyDeviceToHost));
    x->host_valid = PETSC_TRUE;
  }
  return 0;
}
PetscErrorCode vec_to_device(Vec x, bool copy)
{
  if (!have_device()) SETERRQ(PETSC_COMM_SELF, 92, "vector arithmetic needs a B200: there is no CPU fallback");
  if (!x->d_array) CUDA_CHK(cudaMalloc((void **)&x->d_array, sizeof(PetscScalar) * (size_t)PetscMax(x->n, 1)));
  if (copy && !x->dev_valid) {
    if (x->array) CUDA_CHK(cudaMemcpy(x->d_array, x->array, sizeof(PetscScalar) * (size_t)x->n, cudaMemcpyHostToDevice));
    else CUDA_CHK(cudaMemset(x->d_array, 0, sizeof(PetscScalar) * (size_t)x->n));   // never touched on the host: zeros
    x->dev_valid = PETSC_TRUE;
  }
  return 0;
}
}  // namespace

extern "C" PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec *v)
{
  if (n < 0) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "negative vector length");
  Vec x = (Vec)calloc(1, sizeof(struct _p_Vec));
  if (!x) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_MEM, "out of memory");
  x->n = n;
  x->array = NULL;
  x->host_valid = PETSC_TRUE;
  x->dev_valid = PETSC_FALSE;
  *v = x;
  return 0;
}
extern "C" PetscErrorCode VecDuplicate(Vec v, Vec *newv) { return VecCreateSeq(PETSC_COMM_SELF, v->n, newv); }
extern "C" PetscErrorCode VecDestroy(Vec *v)
{
  if (!v || !*v) return 0;
  Vec x = *v;
  if (x->d_array) cudaFree(x->d_array);
  if (x->array) {
    if (g_vecpriv[x].pinned) b200_host_free(x->array); else free(x->array);
    g_vecpriv.erase(x);
  }
  free(x);
  *v = NULL;
  return 0;
}
extern "C" PetscErrorCode VecGetSize(Vec x, PetscInt *n) { *n = x->n; return 0; }
extern "C" PetscErrorCode VecGetLocalSize(Vec x, PetscInt *n) { *n = x->n; return 0; }
extern "C" PetscErrorCode VecGetArray(Vec x, PetscScalar **a)
{
  PetscErrorCode ierr = vec_to_host(x);CHKERRQ(ierr);
  x->dev_valid = PETSC_FALSE;  // the caller may write
  x->state++;
  *a = x->array;
  return 0;
}
extern "C" PetscErrorCode VecRestoreArray(Vec, PetscScalar **a) { if (a) *a = NULL; return 0; }
extern "C" PetscErrorCode VecGetArrayRead(Vec x, const PetscScalar **a)
{
  PetscErrorCode ierr = vec_to_host(x);CHKERRQ(ierr);
  *a = x->array;
  return 0;
}
extern "C" PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **a) { if (a) *a = NULL; return 0; }
extern "C" PetscErrorCode VecB200HasDevice(Vec x, PetscBool *flg) { *flg = (x->dev_valid && have_device()) ? PETSC_TRUE : PETSC_FALSE; return 0; }
extern "C" PetscErrorCode VecB200GetDeviceArrayRead(Vec x, const PetscScalar **d)
{
  PetscErrorCode ierr = vec_to_device(x, true);CHKERRQ(ierr);
  *d = x->d_array;
  return 0;
}
extern "C" PetscErrorCode VecB200GetDeviceArray(Vec x, PetscScalar **d)
{
  PetscErrorCode ierr = vec_to_device(x, true);CHKERRQ(ierr);
  x->host_valid = PETSC_FALSE;
  x->state++;
  *d = x->d_array;
  return 0;
}
extern "C" PetscErrorCode VecB200GetDeviceArrayWrite(Vec x, PetscScalar **d)
{
  PetscErrorCode ierr = vec_to_device(x, false);CHKERRQ(ierr);
  x->dev_valid = PETSC_TRUE;
  x->host_valid = PETSC_FALSE;
  x->state++;
  *d = x->d_array;
  return 0;
}
extern "C" PetscErrorCode VecSet(Vec x, PetscScalar alpha)
{
  if (have_device()) {
    PetscScalar *d;
    PetscErrorCode ierr = VecB200GetDeviceArrayWrite(x, &d);CHKERRQ(ierr);
    B200_CHK(b200_vec_set(d, alpha, x->n, NULL));
    return 0;
  }
  PetscErrorCode ierr = vec_host_alloc(x);CHKERRQ(ierr);
  for (PetscInt i = 0; i < x->n; ++i) x->array[i] = alpha;  // pure assignment, no arithmetic
  x->host_valid = PETSC_TRUE; x->dev_valid = PETSC_FALSE; x->state++;
  return 0;
}
extern "C" PetscErrorCode VecCopy(Vec x, Vec y)
{
  if (x->n != y->n) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "VecCopy: size mismatch");
  if (x == y) return 0;
  // on the device when x is current there, or when y has never been touched on the host (a work
  // vector of a solve: uploading x once is cheaper than giving y a page-locked host array)
  if (have_device() && (x->dev_valid || !y->array)) {
    const PetscScalar *dx;
    PetscScalar *d;
    PetscErrorCode ierr = VecB200GetDeviceArrayRead(x, &dx);CHKERRQ(ierr);
    ierr = VecB200GetDeviceArrayWrite(y, &d);CHKERRQ(ierr);
    B200_CHK(b200_vec_copy(d, dx, x->n, NULL));
    return 0;
  }
  PetscErrorCode ierr = vec_host_alloc(x);CHKERRQ(ierr);
  ierr = vec_host_alloc(y);CHKERRQ(ierr);
  memcpy(y->array, x->array, sizeof(PetscScalar) * (size_t)x->n);
  y->host_valid = PETSC_TRUE; y->dev_valid = PETSC_FALSE; y->state++;
  return 0;
}
// VecSum_Seq [P376]: sequential host loop -- setup code (src/helper.cpp:266) whose rounding the
// matrix's reference-point value depends on, so the order is kept.
extern "C" PetscErrorCode VecSum(Vec x, PetscScalar *sum)
{
  PetscErrorCode ierr = vec_to_host(x);CHKERRQ(ierr);
  PetscScalar s = 0.0;
  for (PetscInt i = 0; i < x->n; ++i) s += x->array[i];
  *sum = s;
  return 0;
}

namespace {
double *g_scalar = nullptr;  // device scalar for reductions
PetscErrorCode reduce_scalar(PetscReal *out)
{
  CUDA_CHK(cudaMemcpy(out, g_scalar, sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}
PetscErrorCode scalar_slot()
{
  if (!g_scalar) CUDA_CHK(cudaMalloc((void **)&g_scalar, 64));
  return 0;
}
}  // namespace

extern "C" PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val)
{
  const PetscScalar *dx, *dy;
  PetscErrorCode ierr;
  ierr = VecB200GetDeviceArrayRead(x, &dx);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayRead(y, &dy);CHKERRQ(ierr);
  ierr = scalar_slot();CHKERRQ(ierr);
  B200_CHK(b200_vec_dot(dx, dy, x->n, g_scalar, NULL));
  ierr = PetscLogFlops(2.0 * x->n - 1);CHKERRQ(ierr);
  return reduce_scalar(val);
}
extern "C" PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val)
{
  const PetscScalar *dx;
  PetscErrorCode ierr;
  ierr = VecB200GetDeviceArrayRead(x, &dx);CHKERRQ(ierr);
  ierr = scalar_slot();CHKERRQ(ierr);
  if (type == NORM_2 || type == NORM_FROBENIUS) B200_CHK(b200_vec_norm2(dx, x->n, g_scalar, NULL));
  else if (type == NORM_INFINITY) B200_CHK(b200_vec_norm_inf(dx, x->n, g_scalar, NULL));
  else SETERRQ(PETSC_COMM_SELF, PETSC_ERR_SUP, "VecNorm: only NORM_2 and NORM_INFINITY");
  return reduce_scalar(val);
}
extern "C" PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x)
{
  const PetscScalar *dx;
  PetscScalar *dy;
  PetscErrorCode ierr;
  ierr = VecB200GetDeviceArrayRead(x, &dx);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArray(y, &dy);CHKERRQ(ierr);
  B200_CHK(b200_vec_axpy(dy, alpha, dx, x->n, NULL));
  return PetscLogFlops(2.0 * x->n);
}
extern "C" PetscErrorCode VecAYPX(Vec y, PetscScalar alpha, Vec x)
{
  const PetscScalar *dx;
  PetscScalar *dy;
  PetscErrorCode ierr;
  ierr = VecB200GetDeviceArrayRead(x, &dx);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArray(y, &dy);CHKERRQ(ierr);
  B200_CHK(b200_vec_aypx(dy, alpha, dx, x->n, NULL));
  return PetscLogFlops(2.0 * x->n);
}
extern "C" PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y)
{
  const PetscScalar *dx, *dy;
  PetscScalar *dw;
  PetscErrorCode ierr;
  ierr = VecB200GetDeviceArrayRead(x, &dx);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayRead(y, &dy);CHKERRQ(ierr);
  if (w == x || w == y) { ierr = VecB200GetDeviceArray(w, &dw);CHKERRQ(ierr); }
  else { ierr = VecB200GetDeviceArrayWrite(w, &dw);CHKERRQ(ierr); }
  B200_CHK(b200_vec_pointwise_mult(dw, dx, dy, x->n, NULL));
  return PetscLogFlops((double)x->n);
}
// VecReciprocal [P376]: x_i <- 1/x_i where x_i != 0 (PCJACOBI setup; host, once per solve)
extern "C" PetscErrorCode VecReciprocal(Vec v)
{
  PetscScalar *a;
  PetscErrorCode ierr = VecGetArray(v, &a);CHKERRQ(ierr);
  for (PetscInt i = 0; i < v->n; ++i) if (a[i] != 0.0) a[i] = 1.0 / a[i];
  return VecRestoreArray(v, &a);
}

// ---------------------------------------------------------------------------------------------
// Mat (SeqAIJ)
// ---------------------------------------------------------------------------------------------
static PetscErrorCode MatSetValues_SeqAIJ(Mat, PetscInt, const PetscInt[], PetscInt, const PetscInt[], const PetscScalar[], InsertMode);
static PetscErrorCode MatGetDiagonal_SeqAIJ(Mat, Vec);
static PetscErrorCode MatZeroRowsColumns_SeqAIJ(Mat, PetscInt, const PetscInt[], PetscScalar, Vec, Vec);
static PetscErrorCode MatScale_SeqAIJ(Mat, PetscScalar);

// MatCreateSeqAIJ + MatSeqAIJSetPreallocation [P376]: row i owns imax[i] consecutive slots.
extern "C" PetscErrorCode MatCreateSeqAIJ(MPI_Comm, PetscInt m, PetscInt n, PetscInt nz, const PetscInt nnz[], Mat *newA)
{
  if (m < 0 || n < 0) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "negative matrix size");
  if (nz == PETSC_DEFAULT || nz == PETSC_DECIDE) nz = 5;
  Mat A = (Mat)calloc(1, sizeof(struct _p_Mat));
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)calloc(1, sizeof(Mat_SeqAIJ));
  if (!A || !a) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_MEM, "out of memory");
  A->data = a;
  A->type_name = MATSEQAIJ;
  A->rmap = (PetscLayout)calloc(1, sizeof(struct _n_PetscLayout));
  A->cmap = (PetscLayout)calloc(1, sizeof(struct _n_PetscLayout));
  A->rmap->n = A->rmap->N = m; A->rmap->rend = m;
  A->cmap->n = A->cmap->N = n; A->cmap->rend = n;
  // the operator table: these are the symbols the reference replaces
  A->ops->mult = MatMult_SeqAIJ;
  A->ops->multadd = MatMultAdd_SeqAIJ;
  A->ops->multtranspose = MatMultTranspose_SeqAIJ;
  A->ops->multtransposeadd = MatMultTransposeAdd_SeqAIJ;
  A->ops->assemblyend = MatAssemblyEnd_SeqAIJ;
  A->ops->destroy = MatDestroy_SeqAIJ;
  A->ops->getdiagonal = MatGetDiagonal_SeqAIJ;
  A->ops->setvalues = MatSetValues_SeqAIJ;
  A->ops->zerorowscolumns = MatZeroRowsColumns_SeqAIJ;
  A->ops->scale = MatScale_SeqAIJ;
  a->imax = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)PetscMax(m, 1));
  a->ilen = (PetscInt *)calloc((size_t)PetscMax(m, 1), sizeof(PetscInt));
  a->i = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)m + 1));
  long long tot = 0;
  for (PetscInt r = 0; r < m; ++r) {
    PetscInt k = nnz ? nnz[r] : nz;
    if (k < 0) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "nnz cannot be less than 0: local row %d value %d", r, k);
    if (k > n) k = n;
    a->imax[r] = k;
    a->i[r] = (PetscInt)tot;
    tot += k;
  }
  if (tot > 2147483647LL) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "preallocation exceeds int32 indices");
  a->i[m] = (PetscInt)tot;
  a->maxnz = (PetscInt)tot;
  a->j = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)PetscMax(tot, 1));
  a->a = (MatScalar *)calloc((size_t)PetscMax(tot, 1), sizeof(MatScalar));
  a->nz = 0;
  a->singlemalloc = PETSC_FALSE;
  a->roworiented = PETSC_TRUE;
  a->nounused = 0;
  *newA = A;
  return 0;
}

// MatSetValues_SeqAIJ [P376]: per value, find the column in the sorted used part of the row
// (INSERT overwrites, ADD accumulates); otherwise shift the tail up and insert; negative row or
// column indices are ignored (this is what drops the ghost neighbours, src/helper.cpp:233-236).
// When the reserved slots of a row are full the whole CSR is re-laid-out with room for more
// (PETSc: MatSeqXAIJReallocateAIJ; counted in a->reallocs).
static PetscErrorCode grow_row(Mat A, PetscInt row, PetscInt extra)
{
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  const PetscInt m = A->rmap->n;
  const long long newmax = (long long)a->maxnz + extra;
  if (newmax > 2147483647LL) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "matrix exceeds int32 indices");
  PetscInt  *nj = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)newmax);
  MatScalar *na = (MatScalar *)calloc((size_t)newmax, sizeof(MatScalar));
  if (!nj || !na) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_MEM, "out of memory");
  const PetscInt cut = a->i[row] + a->ilen[row];
  memcpy(nj, a->j, sizeof(PetscInt) * (size_t)cut);
  memcpy(na, a->a, sizeof(MatScalar) * (size_t)cut);
  memcpy(nj + cut + extra, a->j + cut, sizeof(PetscInt) * (size_t)(a->maxnz - cut));
  memcpy(na + cut + extra, a->a + cut, sizeof(MatScalar) * (size_t)(a->maxnz - cut));
  for (PetscInt r = row + 1; r <= m; ++r) a->i[r] += extra;
  a->imax[row] += extra;
  free(a->j); free(a->a);
  a->j = nj; a->a = na;
  a->maxnz = (PetscInt)newmax;
  a->reallocs++;
  return 0;
}

static PetscErrorCode MatSetValues_SeqAIJ(Mat A, PetscInt m, const PetscInt im[], PetscInt n, const PetscInt in[], const PetscScalar v[], InsertMode is)
{
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  for (PetscInt k = 0; k < m; ++k) {
    const PetscInt row = im[k];
    if (row < 0) continue;
    if (row >= A->rmap->n) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "Row too large: row %d max %d", row, A->rmap->n - 1);
    for (PetscInt l = 0; l < n; ++l) {
      const PetscInt col = in[l];
      if (col < 0) continue;
      if (col >= A->cmap->n) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "Column too large: col %d max %d", col, A->cmap->n - 1);
      const PetscScalar value = v[l + k * n];
      PetscInt  *rp = a->j + a->i[row];
      MatScalar *ap = a->a + a->i[row];
      PetscInt   nrow = a->ilen[row], lo = 0, hi = nrow;
      while (hi - lo > 5) {  // bisection, then a short linear scan (PETSc's search shape)
        PetscInt t = (lo + hi) / 2;
        if (rp[t] > col) hi = t; else lo = t;
      }
      PetscInt pos = lo;
      bool found = false;
      for (; pos < hi; ++pos) {
        if (rp[pos] > col) break;
        if (rp[pos] == col) { found = true; break; }
      }
      if (found) {
        if (is == ADD_VALUES) ap[pos] += value; else ap[pos] = value;
        continue;
      }
      if (nrow >= a->imax[row]) {
        PetscErrorCode ierr = grow_row(A, row, PetscMax(10, a->imax[row]));CHKERRQ(ierr);
        rp = a->j + a->i[row];
        ap = a->a + a->i[row];
      }
      for (PetscInt t = nrow - 1; t >= pos; --t) { rp[t + 1] = rp[t]; ap[t + 1] = ap[t]; }
      rp[pos] = col;
      ap[pos] = value;
      a->ilen[row] = nrow + 1;
    }
  }
  A->state++;
  A->assembled = PETSC_FALSE;
  return 0;
}

extern "C" PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode addv)
{
  return (*A->ops->setvalues)(A, m, idxm, n, idxn, v, addv);
}
extern "C" PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
extern "C" PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType type)
{
  PetscErrorCode ierr = (*A->ops->assemblyend)(A, type);CHKERRQ(ierr);
  if (type == MAT_FINAL_ASSEMBLY) A->assembled = PETSC_TRUE;
  A->state++;
  return 0;
}
#define CHECK_ASSEMBLED(A) do { if (!(A)->assembled) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "Not for unassembled matrix"); } while (0)
extern "C" PetscErrorCode MatMult(Mat A, Vec x, Vec y)
{
  CHECK_ASSEMBLED(A);
  if (x == y) SETERRQ(PETSC_COMM_SELF, 61, "x and y must be different vectors");
  if (A->cmap->n != x->n || A->rmap->n != y->n) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "Mat/Vec size mismatch");
  return (*A->ops->mult)(A, x, y);
}
extern "C" PetscErrorCode MatMultAdd(Mat A, Vec v1, Vec v2, Vec v3)
{
  CHECK_ASSEMBLED(A);
  if (v1 == v3) SETERRQ(PETSC_COMM_SELF, 61, "v1 and v3 must be different vectors");
  if (A->cmap->n != v1->n || A->rmap->n != v2->n || A->rmap->n != v3->n) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "Mat/Vec size mismatch");
  return (*A->ops->multadd)(A, v1, v2, v3);
}
extern "C" PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y)
{
  CHECK_ASSEMBLED(A);
  if (x == y) SETERRQ(PETSC_COMM_SELF, 61, "x and y must be different vectors");
  if (A->rmap->n != x->n || A->cmap->n != y->n) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "Mat/Vec size mismatch");
  return (*A->ops->multtranspose)(A, x, y);
}
extern "C" PetscErrorCode MatMultTransposeAdd(Mat A, Vec v1, Vec v2, Vec v3)
{
  CHECK_ASSEMBLED(A);
  if (v1 == v3) SETERRQ(PETSC_COMM_SELF, 61, "v1 and v3 must be different vectors");
  if (A->rmap->n != v1->n || A->cmap->n != v2->n || A->cmap->n != v3->n) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "Mat/Vec size mismatch");
  return (*A->ops->multtransposeadd)(A, v1, v2, v3);
}
extern "C" PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n) { if (m) *m = A->rmap->N; if (n) *n = A->cmap->N; return 0; }
extern "C" PetscErrorCode MatGetLocalSize(Mat A, PetscInt *m, PetscInt *n) { if (m) *m = A->rmap->n; if (n) *n = A->cmap->n; return 0; }
extern "C" PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left)
{
  PetscErrorCode ierr;
  if (right) { ierr = VecCreateSeq(PETSC_COMM_SELF, A->cmap->n, right);CHKERRQ(ierr); }
  if (left) { ierr = VecCreateSeq(PETSC_COMM_SELF, A->rmap->n, left);CHKERRQ(ierr); }
  return 0;
}
extern "C" PetscErrorCode MatGetDiagonal(Mat A, Vec v) { CHECK_ASSEMBLED(A); return (*A->ops->getdiagonal)(A, v); }
extern "C" PetscErrorCode MatZeroRowsColumns(Mat A, PetscInt n, const PetscInt rows[], PetscScalar diag, Vec x, Vec b)
{
  CHECK_ASSEMBLED(A);
  return (*A->ops->zerorowscolumns)(A, n, rows, diag, x, b);
}
extern "C" PetscErrorCode MatScale(Mat A, PetscScalar s) { return (*A->ops->scale)(A, s); }
extern "C" PetscErrorCode MatDestroy(Mat *A)
{
  if (!A || !*A) return 0;
  PetscErrorCode ierr = (*(*A)->ops->destroy)(*A);CHKERRQ(ierr);
  free((*A)->rmap); free((*A)->cmap);
  free(*A);
  *A = NULL;
  return 0;
}

// --- private helpers with PETSc's names ---------------------------------------------------------
extern "C" PetscErrorCode MatMarkDiagonal_SeqAIJ(Mat A)
{
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  const PetscInt m = A->rmap->n;
  if (!a->diag) a->diag = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)PetscMax(m, 1));
  for (PetscInt i = 0; i < m; ++i) {
    a->diag[i] = a->i[i + 1];  // PETSc: "missing" marker is the end of the row
    for (PetscInt k = a->i[i]; k < a->i[i + 1]; ++k) if (a->j[k] == i) { a->diag[i] = k; break; }
  }
  return 0;
}
extern "C" PetscErrorCode MatCheckCompressedRow(Mat, PetscInt nrows, Mat_CompressedRow *c, PetscInt *ai, PetscInt mbs, PetscReal ratio)
{
  free(c->i); free(c->rindex);
  c->i = NULL; c->rindex = NULL;
  nrows = mbs - nrows;  // number of zero rows
  if (nrows < ratio * mbs) { c->use = PETSC_FALSE; c->nrows = 0; return 0; }
  c->use = PETSC_TRUE;
  nrows  = mbs - nrows;
  c->i = (PetscInt *)malloc(sizeof(PetscInt) * ((size_t)nrows + 1));
  c->rindex = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)PetscMax(nrows, 1));
  PetscInt row = 0;
  c->i[0] = 0;
  for (PetscInt i = 0; i < mbs; ++i) {
    if (ai[i + 1] - ai[i] == 0) continue;
    c->i[row + 1] = ai[i + 1];
    c->rindex[row++] = i;
  }
  c->nrows = nrows;
  return 0;
}
extern "C" PetscErrorCode MatAssemblyEnd_SeqAIJ_Inode(Mat, MatAssemblyType) { return 0; }
extern "C" PetscErrorCode MatSeqAIJInvalidateDiagonal(Mat) { return 0; }
extern "C" PetscErrorCode MatSeqXAIJFreeAIJ(Mat, MatScalar **a, PetscInt **j, PetscInt **i)
{
  free(*a); free(*j); free(*i);
  *a = NULL; *j = NULL; *i = NULL;
  return 0;
}

static PetscErrorCode MatGetDiagonal_SeqAIJ(Mat A, Vec v)
{
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  PetscScalar *x;
  PetscErrorCode ierr = VecGetArray(v, &x);CHKERRQ(ierr);
  for (PetscInt i = 0; i < A->rmap->n; ++i) {
    x[i] = 0.0;
    for (PetscInt k = a->i[i]; k < a->i[i + 1]; ++k) if (a->j[k] == i) { x[i] = a->a[k]; break; }
  }
  return VecRestoreArray(v, &x);
}

// MatZeroRowsColumns_SeqAIJ [P376] (called by src/helper.cpp:272): the pattern is kept; listed
// rows become zero with `diag` on the diagonal, listed columns are zeroed in the other rows with
// b[i] -= a_ij x_j, and b[row] = diag * x[row].
static PetscErrorCode MatZeroRowsColumns_SeqAIJ(Mat A, PetscInt N, const PetscInt rows[], PetscScalar diag, Vec x, Vec b)
{
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  const PetscInt m = A->rmap->n;
  const PetscScalar *xx = NULL;
  PetscScalar *bb = NULL;
  PetscErrorCode ierr;
  const bool vecs = x && b;
  if (vecs) { ierr = VecGetArrayRead(x, &xx);CHKERRQ(ierr); ierr = VecGetArray(b, &bb);CHKERRQ(ierr); }
  std::vector<char> zeroed((size_t)PetscMax(m, 1), 0);
  for (PetscInt i = 0; i < N; ++i) {
    if (rows[i] < 0 || rows[i] >= m) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "row %d out of range", rows[i]);
    zeroed[rows[i]] = 1;
    memset(a->a + a->i[rows[i]], 0, sizeof(MatScalar) * (size_t)(a->i[rows[i] + 1] - a->i[rows[i]]));
  }
  for (PetscInt i = 0; i < m; ++i) {
    if (!zeroed[i]) {
      for (PetscInt k = a->i[i]; k < a->i[i + 1]; ++k)
        if (a->j[k] < m && zeroed[a->j[k]]) {
          if (vecs) bb[i] -= a->a[k] * xx[a->j[k]];
          a->a[k] = 0.0;
        }
    } else if (vecs) bb[i] = diag * xx[i];
  }
  if (diag != 0.0) {
    ierr = MatMarkDiagonal_SeqAIJ(A);CHKERRQ(ierr);
    for (PetscInt i = 0; i < N; ++i) {
      if (a->diag[rows[i]] >= a->i[rows[i] + 1]) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "Matrix is missing diagonal entry in row %d", rows[i]);
      a->a[a->diag[rows[i]]] = diag;
    }
  }
  if (vecs) { ierr = VecRestoreArrayRead(x, &xx);CHKERRQ(ierr); ierr = VecRestoreArray(b, &bb);CHKERRQ(ierr); }
  A->state++;
  return MatAssemblyEnd_SeqAIJ(A, MAT_FINAL_ASSEMBLY);
}

// MatScale_SeqAIJ: in-place value change; the reference's pointer-keyed residency would miss it
static PetscErrorCode MatScale_SeqAIJ(Mat A, PetscScalar s)
{
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  for (PetscInt k = 0; k < a->nz; ++k) a->a[k] *= s;
  A->state++;
  return PetscLogFlops((double)a->nz);
}

// An assembled SeqAIJ matrix from a finished CSR (copied): what MatCreateSeqAIJWithArrays gives in
// PETSc, except that the Mat owns its arrays.  Used for the multigrid level operators.
extern "C" PetscErrorCode MatCreateSeqAIJFromCSRB200(PetscInt m, PetscInt n, const PetscInt i[], const PetscInt j[], const PetscScalar v[], Mat *newA)
{
  if (m < 0 || n < 0 || !i || i[0] != 0) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "bad CSR");
  std::vector<PetscInt> len((size_t)PetscMax(m, 1), 0);
  for (PetscInt r = 0; r < m; ++r) {
    len[r] = i[r + 1] - i[r];
    if (len[r] < 0) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "row pointers decrease at row %d", r);
    for (PetscInt k = i[r]; k < i[r + 1]; ++k) {
      if (j[k] < 0 || j[k] >= n) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "column %d out of range in row %d", j[k], r);
      if (k > i[r] && j[k] <= j[k - 1]) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "columns of row %d are not strictly ascending", r);
    }
  }
  PetscErrorCode ierr = MatCreateSeqAIJ(PETSC_COMM_SELF, m, n, 0, len.data(), newA);CHKERRQ(ierr);
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)(*newA)->data;
  if (m && i[m]) {
    memcpy(a->j, j, sizeof(PetscInt) * (size_t)i[m]);
    memcpy(a->a, v, sizeof(MatScalar) * (size_t)i[m]);
  }
  for (PetscInt r = 0; r < m; ++r) a->ilen[r] = len[r];
  return MatAssemblyEnd(*newA, MAT_FINAL_ASSEMBLY);
}

extern "C" PetscErrorCode MatSeqAIJGetCSRB200(Mat A, PetscInt *m, PetscInt *n, PetscInt *nz, const PetscInt **i, const PetscInt **j, const PetscScalar **v)
{
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  if (m) *m = A->rmap->n;
  if (n) *n = A->cmap->n;
  if (nz) *nz = a->nz;
  if (i) *i = a->i;
  if (j) *j = a->j;
  if (v) *v = a->a;
  return 0;
}
extern "C" PetscErrorCode MatSeqAIJGetInfoB200(Mat A, PetscInt *nonzerorowcnt, PetscInt *rmax, PetscBool *cprow, PetscInt *cprow_nrows, PetscInt *fshift)
{
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  if (nonzerorowcnt) *nonzerorowcnt = a->nonzerorowcnt;
  if (rmax) *rmax = a->rmax;
  if (cprow) *cprow = a->compressedrow.use;
  if (cprow_nrows) *cprow_nrows = a->compressedrow.nrows;
  if (fshift) *fshift = (PetscInt)A->info.nz_unneeded;
  return 0;
}
