/* ============================================================================
 * cnld_b200_h2lib.h -- the H2Lib-named part of libcnld_b200.so's C ABI.
 *
 * These are the symbols cnl-dyna's Cython layer binds (`cdef extern` blocks in
 * /root/reference/cnld/h2lib/_*.pxd), with identical prototypes, so that
 * cnld/h2lib/*.pyx can be re-pointed at libcnld_b200.so instead of libh2.a
 * (setup.py:15-16).  Struct members that Cython (or Python through numpy views)
 * dereferences keep H2Lib's names, types and order; private members follow them.
 * All `a`, `v`, `x`, ... buffers are HOST memory (numpy views alias them,
 * cnld/h2lib/amatrix.pyx:34-37); the arithmetic entry points move the data
 * to the GPU, run the sm_100a kernels and copy the result back.  For a whole
 * frequency sweep use the device-resident entry points in cnld_b200.h instead.
 *
 * real = double, field = double _Complex, uint = unsigned (settings.h:62,89,163).
 * ========================================================================== */
#ifndef CNLD_B200_H2LIB_H
#define CNLD_B200_H2LIB_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
typedef struct { double re, im; } field;
#else
typedef double _Complex field;
#endif
typedef double real;
typedef unsigned int uint;
typedef field *pfield;
typedef real *preal;

/* ---- basic (cnld/h2lib/_basic.pxd:20-25; include/basic.h) ------------------ */
void init_h2lib(int *argc, char ***argv);
void uninit_h2lib(void);
void freemem(void *ptr);
/* macros in H2Lib (include/basic.h:181,298,302), used by the Cython parametrizations
 * (cnld/h2lib/macrosurface3d.pyx:83-172) */
#include <math.h>
#ifndef REAL_ABS
#define REAL_ABS(x) fabs(x)
#define REAL_NORM2(x, y) sqrt((x) * (x) + (y) * (y))
#define REAL_NORM3(x, y, z) sqrt((x) * (x) + (y) * (y) + (z) * (z))
#endif

/* ---- avector (cnld/h2lib/_avector.pxd:8-23; include/avector.h:39-48) ------- */
typedef struct _avector {
  field *v;
  uint dim;
  void *owner;
} avector, *pavector;
typedef const avector *pcavector;
pavector new_avector(uint dim);
pavector new_zero_avector(uint dim);
void del_avector(pavector v);
void clear_avector(pavector v);
void copy_avector(pcavector v, pavector w);
void random_avector(pavector v);

/* ---- amatrix (cnld/h2lib/_amatrix.pxd:8-34; include/amatrix.h:42-57) ------- */
typedef struct _amatrix {
  field *a;   /* column-major, a[i + j*ld] */
  uint ld;
  uint rows;
  uint cols;
  void *owner;
} amatrix, *pamatrix;
typedef const amatrix *pcamatrix;
pamatrix new_amatrix(uint rows, uint cols);
pamatrix new_zero_amatrix(uint rows, uint cols);
void del_amatrix(pamatrix a);
pamatrix clone_amatrix(pcamatrix src);
void scale_amatrix(field alpha, pamatrix a);
void conjugate_amatrix(pamatrix a);
void addeval_amatrix_avector(field alpha, pcamatrix a, pcavector src, pavector trg);
void add_amatrix(field alpha, bool atrans, pcamatrix a, pamatrix b);
size_t getsize_amatrix(pcamatrix a);
void addmul_amatrix(field alpha, bool atrans, pcamatrix a, bool btrans, pcamatrix b, pamatrix c);

/* ---- sparsematrix (cnld/h2lib/_sparsematrix.pxd:12-30; sparsematrix.h:41-57) */
typedef struct _sparsematrix {
  uint rows;
  uint cols;
  uint nz;
  uint *row;    /* CRS row starts, rows+1 entries */
  uint *col;
  pfield coeff;
} sparsematrix, *psparsematrix;
typedef const sparsematrix *pcsparsematrix;
psparsematrix new_raw_sparsematrix(uint rows, uint cols, uint nz);
psparsematrix new_identity_sparsematrix(uint rows, uint cols);
void del_sparsematrix(psparsematrix a);
field addentry_sparsematrix(psparsematrix a, uint row, uint col, field x);
void setentry_sparsematrix(psparsematrix a, uint row, uint col, field x);
size_t getsize_sparsematrix(pcsparsematrix a);
void sort_sparsematrix(psparsematrix a);
void clear_sparsematrix(psparsematrix a);
void add_sparsematrix_amatrix(field alpha, bool atrans, pcsparsematrix a, pamatrix b);
void addeval_sparsematrix_avector(field alpha, pcsparsematrix a, pcavector x, pavector y);

/* ---- surface3d (cnld/h2lib/_surface3d.pxd:9-34; include/surface3d.h:45-71) -- */
typedef struct _surface3d {
  uint vertices;
  uint edges;
  uint triangles;
  real (*x)[3];
  uint (*e)[2];
  uint (*t)[3];
  uint (*s)[3];   /* s[i][j] lies opposite t[i][j] */
  real (*n)[3];
  preal g;        /* Gram determinant = 2*area */
  real hmin;
  real hmax;
} surface3d, *psurface3d;
typedef const surface3d *pcsurface3d;
psurface3d new_surface3d(uint vertices, uint edges, uint triangles);
void del_surface3d(psurface3d gr);
void prepare_surface3d(psurface3d gr);
psurface3d merge_surface3d(pcsurface3d gr1, pcsurface3d gr2);
psurface3d refine_red_surface3d(psurface3d gr);
void translate_surface3d(psurface3d gr, real *t);
uint check_surface3d(pcsurface3d gr);
bool isclosed_surface3d(pcsurface3d gr);
bool isoriented_surface3d(pcsurface3d gr);

/* ---- macrosurface3d (_macrosurface3d.pxd:8-26; macrosurface3d.h:48-101,187-199) */
typedef struct _macrosurface3d {
  uint vertices;
  uint edges;
  uint triangles;
  real (*x)[3];
  uint (*e)[2];
  uint (*t)[3];
  uint (*s)[3];
  void (*phi)(uint i, real xr1, real xr2, void *phidata, real xt[3]);
  void *phidata;
} macrosurface3d, *pmacrosurface3d;
typedef const macrosurface3d *pcmacrosurface3d;
pmacrosurface3d new_macrosurface3d(uint vertices, uint edges, uint triangles);
void del_macrosurface3d(pmacrosurface3d mg);
psurface3d build_from_macrosurface3d_surface3d(pcmacrosurface3d mg, uint split);
pmacrosurface3d new_sphere_macrosurface3d(void);
/* the parametrizations cnl-dyna implements in Cython (macrosurface3d.pyx:83-172),
 * provided natively so a ctypes host does not need callbacks */
void cnld_b200_phi_affine(uint i, real xr1, real xr2, void *phidata, real xt[3]);
void cnld_b200_phi_circle(uint i, real xr1, real xr2, void *phidata, real xt[3]);
void cnld_b200_phi_sphere(uint i, real xr1, real xr2, void *phidata, real xt[3]);

/* ---- bem3d / helmholtzbem3d (_bem3d.pxd:10-37, _helmholtzbem3d.pxd:11-14) --- */
typedef enum {
  BASIS_NONE_BEM3D = 0,
  BASIS_CONSTANT_BEM3D = 'c',
  BASIS_LINEAR_BEM3D = 'l'
} basisfunctionbem3d;
typedef struct _bem3d {
  pcsurface3d gr;
  void *sq;                       /* quadrature tables (device + host), private */
  basisfunctionbem3d row_basis;
  basisfunctionbem3d col_basis;
  real *mass;
  field alpha;
  field k;                        /* wavenumber (bem3d.h:338) */
  field kernel_const;             /* 1/(4 pi) (helmholtzbem3d.h:35) */
  void *priv;                     /* cnld_ctx of the device-side state */
  uint q_regular, q_singular;
  real accur;                     /* ACA accuracy set by setup_hmatrix_aprx_*_bem3d */
} bem3d, *pbem3d;
typedef const bem3d *pcbem3d;
pbem3d new_bem3d(pcsurface3d gr, basisfunctionbem3d row_basis, basisfunctionbem3d col_basis);
void del_bem3d(pbem3d bem);
pbem3d new_slp_helmholtz_bem3d(field k, pcsurface3d gr, uint q_regular, uint q_singular,
                               basisfunctionbem3d row_basis, basisfunctionbem3d col_basis);
/* bound by _helmholtzbem3d.pxd:13 (helmholtzbem3d.h:113) but never constructed by cnl-dyna: no kernel here,
 * the call reports that and returns NULL ("what links and what does not") */
pbem3d new_dlp_helmholtz_bem3d(field k, pcsurface3d gr, uint q_regular, uint q_singular,
                               basisfunctionbem3d row_basis, basisfunctionbem3d col_basis, field alpha);
void assemble_bem3d_amatrix(pbem3d bem, pamatrix G);

/* ---- cluster / block trees (_cluster.pxd:7-23, _block.pxd:8-31; cluster.h:40-70, block.h:48-69) */
typedef struct _cluster cluster, *pcluster;
typedef const cluster *pccluster;
struct _cluster {
  uint size;
  uint *idx;        /* sons' idx point into the father's array (root owns it) */
  uint sons;
  pcluster *son;
  uint dim;
  real *bmin;
  real *bmax;
  uint desc;
  uint type;
};
pcluster new_cluster(uint size, uint *idx, uint sons, uint dim);
void del_cluster(pcluster t);
typedef struct _block block, *pblock;
typedef const block *pcblock;
struct _block {
  pcluster rc;
  pcluster cc;
  bool a;           /* admissible */
  pblock *son;      /* son[i + j*rsons] */
  uint rsons;
  uint csons;
  uint desc;
};
typedef bool (*admissible)(pcluster rc, pcluster cc, void *data);
pblock new_block(pcluster rc, pcluster cc, bool a, uint rsons, uint csons);
void del_block(pblock b);
pblock build_nonstrict_block(pcluster rc, pcluster cc, void *data, admissible admis);
pblock build_strict_block(pcluster rc, pcluster cc, void *data, admissible admis);
bool admissible_2_cluster(pcluster rc, pcluster cc, void *data);       /* data = real *eta */
bool admissible_max_cluster(pcluster rc, pcluster cc, void *data);
bool admissible_sphere_cluster(pcluster rc, pcluster cc, void *data);
bool admissible_2_min_cluster(pcluster rc, pcluster cc, void *data);
pcluster build_bem3d_cluster(pcbem3d bem, uint clf, basisfunctionbem3d basis);

/* ---- rkmatrix / hmatrix (_rkmatrix.pxd:7-19, _hmatrix.pxd:12-46; rkmatrix.h:36-44, hmatrix.h:49-72) */
typedef struct _rkmatrix {
  amatrix A;        /* rows x k */
  amatrix B;        /* cols x k; block = A B^* */
  uint k;
} rkmatrix, *prkmatrix;
typedef const rkmatrix *pcrkmatrix;
prkmatrix new_rkmatrix(uint rows, uint cols, uint k);
void del_rkmatrix(prkmatrix r);
typedef struct _hmatrix hmatrix, *phmatrix;
typedef const hmatrix *pchmatrix;
struct _hmatrix {
  pccluster rc;
  pccluster cc;
  prkmatrix r;      /* admissible leaf */
  pamatrix f;       /* inadmissible leaf */
  phmatrix *son;    /* son[i + j*rsons] */
  uint rsons;
  uint csons;
  uint refs;
  uint desc;
};
phmatrix new_hmatrix(pccluster rc, pccluster cc);
void del_hmatrix(phmatrix hm);
void clear_hmatrix(phmatrix hm);
void copy_hmatrix(pchmatrix src, phmatrix trg);
phmatrix clone_hmatrix(pchmatrix src);
phmatrix clonestructure_hmatrix(pchmatrix src);
size_t getsize_hmatrix(pchmatrix hm);
uint getrows_hmatrix(pchmatrix hm);
uint getcols_hmatrix(pchmatrix hm);
phmatrix build_from_block_hmatrix(pcblock b, uint k);
void addeval_hmatrix_avector(field alpha, pchmatrix hm, pcavector x, pavector y);
void addevalsymm_hmatrix_avector(field alpha, pchmatrix hm, pcavector x, pavector y);
void copy_sparsematrix_hmatrix(psparsematrix sp, phmatrix hm);
void identity_hmatrix(phmatrix hm);
void add_hmatrix_amatrix(field alpha, bool atrans, pchmatrix a, pamatrix b);   /* harith.h:88 */
void scale_hmatrix(field alpha, phmatrix hm);   /* not in H2Lib's pxd: exact replacement for the
                                                   truncated identity product of HFormat._smul */
void setup_hmatrix_aprx_aca_bem3d(pbem3d bem, pccluster rc, pccluster cc, pcblock tree, real accur);
void setup_hmatrix_aprx_paca_bem3d(pbem3d bem, pccluster rc, pccluster cc, pcblock tree, real accur);
void setup_hmatrix_aprx_hca_bem3d(pbem3d bem, pccluster rc, pccluster cc, pcblock tree, uint m, real accur);
void setup_hmatrix_aprx_inter_row_bem3d(pbem3d bem, pccluster rc, pccluster cc, pcblock tree, uint m);
void assemble_bem3d_hmatrix(pbem3d bem, pblock b, phmatrix G);

/* ---- truncation (_truncation.pxd:8-22) -- carried for call compatibility only ------------- */
typedef struct _truncmode {
  bool frobenius;
  bool absolute;
  bool blocks;
  real zeta_level;
  real zeta_age;
} truncmode, *ptruncmode;
typedef const truncmode *pctruncmode;
ptruncmode new_truncmode(void);
void del_truncmode(ptruncmode tm);
ptruncmode new_releucl_truncmode(void);

/* ---- krylovsolvers (_krylovsolvers.pxd:10-13; krylovsolvers.h:283) -------------------------- */
/* conjugate gradients for Hermitian positive definite operators; the matvec runs on the GPU */
uint solve_cg_amatrix_avector(pcamatrix A, pcavector b, pavector x, real eps, uint maxiter);
uint solve_cg_hmatrix_avector(pchmatrix A, pcavector b, pavector x, real eps, uint maxiter);
uint solve_gmres_amatrix_avector(pcamatrix A, pcavector b, pavector x, real eps, uint maxiter, uint kmax);
uint solve_gmres_hmatrix_avector(pchmatrix A, pcavector b, pavector x, real eps, uint maxiter, uint kmax);

/* ---- factorizations (_factorizations.pxd:13-19; factorizations.h:173-233) --- */
uint lrdecomp_amatrix(pamatrix a);
void lrsolve_amatrix_avector(bool atrans, pcamatrix a, pavector x);
void triangularsolve_amatrix_avector(bool alower, bool aunit, bool atrans, pcamatrix a, pavector x);
/* Cholesky A = L L^H (factorizations.h:281, :315): lower triangle in place; returns 0 or the failing pivot + 1 */
uint choldecomp_amatrix(pamatrix a);
void cholsolve_amatrix_avector(pcamatrix a, pavector x);

/* ---- matrixnorms (_matrixnorms.pxd:11-19; matrixnorms.h): spectral norm of a difference by 20 power
 * iteration steps, every operator applied on the GPU; the *_lr_* variants compare with the product L R of
 * a factorisation held in one matrix --------------------------------------------------------------- */
real norm2diff_amatrix_sparsematrix(pcsparsematrix a, pcamatrix b);
real norm2diff_amatrix_hmatrix(pchmatrix a, pcamatrix b);
real norm2diff_sparsematrix_hmatrix(pchmatrix a, pcsparsematrix b);
real norm2diff_lr_amatrix(pcamatrix A, pcamatrix LR);
real norm2diff_lr_hmatrix(pchmatrix A, pchmatrix LR);
real norm2diff_lr_amatrix_hmatrix(pcamatrix A, pchmatrix LR);
real norm2diff_lr_hmatrix_amatrix(pchmatrix A, pcamatrix LR);
real norm2diff_lr_sparsematrix_amatrix(pcsparsematrix A, pcamatrix LR);
real norm2diff_lr_sparsematrix_hmatrix(pcsparsematrix A, pchmatrix LR);

/* ---- harith (_harith.pxd:13-23; harith.h:88-630).  Dense-expansion H-arithmetic on the GPU: operands are
 * expanded to dense device matrices, the operation runs on the sweep's dense kernels (ZGEMM, unpivoted LU,
 * Cholesky) and an H-matrix result is re-compressed into the target's block structure (partial-pivot ACA
 * of the explicit block to `eps`; the truncation mode is accepted and ignored).  Factorisations stay dense
 * on the device, attached to the hmatrix they were computed from.  Limited to sizes whose dense expansion
 * fits the device (tens of thousands of DOFs on a B200); larger arrays take the GMRES path. ------------ */
void lrdecomp_hmatrix(phmatrix a, pctruncmode tm, real eps);
void lrsolve_hmatrix_avector(bool atrans, pchmatrix a, pavector x);
void lreval_hmatrix_avector(bool atrans, pchmatrix a, pavector x);
void choldecomp_hmatrix(phmatrix a, pctruncmode tm, real eps);
void cholsolve_hmatrix_avector(pchmatrix a, pavector x);
void choleval_hmatrix_avector(pchmatrix a, pavector x);
void addmul_hmatrix(field alpha, bool xtrans, pchmatrix x, bool ytrans, pchmatrix y, pctruncmode tm, real eps,
                    phmatrix z);
void add_hmatrix(field alpha, pchmatrix a, pctruncmode tm, real eps, phmatrix b);
void add_amatrix_hmatrix(field alpha, bool atrans, pcamatrix a, pctruncmode tm, real eps, phmatrix b);
void triangularsolve_hmatrix_avector(bool alower, bool aunit, bool atrans, pchmatrix a, pavector x);

#ifdef __cplusplus
}
#endif
#endif /* CNLD_B200_H2LIB_H */
