/* petscvec.h -- forwards to the PETSc 3.7.6 API slice in b200_petsc.h (see that file). */
#include "b200_petsc.h"
