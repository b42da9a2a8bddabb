/* Shim so that `cdef extern from 'rkmatrix.h'` in cnl-dyna's cnld/h2lib/_rkmatrix.pxd resolves against
 * libcnld_b200 instead of H2Lib: put include/h2lib_compat first on the include path. */
#include "../cnld_b200_h2lib.h"
