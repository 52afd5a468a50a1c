/* <petscksp.h> as the reference's unmodified sources include it (ref: include/header.h:12, include/mesh.h:11,
 * include/solver.h:14): the PETSc surface served by the B200 engine.  See petsc_b200.h. */
#include "petsc_b200.h"
