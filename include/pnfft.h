/* pnfft-b200: C ABI of the B200-native PNFFT window-convolution path.
 *
 * This header declares the same entry points, handle types and flag VALUES as the reference's
 * public header (reference api/pnfft.h:51-288 prototypes, :302-418 constants), so a caller
 * compiled against PNFFT links against libpnfft_b200.so unchanged.  Each declaration below cites
 * the reference definition it replaces.  Two precisions are exported from one library:
 *   pnfft_*  : R = double, C = double[2]      pnfftf_* : R = float, C = float[2]
 * (the long double instantiation, reference api/pnfft.h:288, has no GPU equivalent).
 *
 * All user-visible arrays (x, f, grad_f, f_hat) may be HOST pointers, as in the reference, or
 * DEVICE pointers: the library detects which (cudaPointerGetAttributes) and skips the copies for
 * device-resident data.  Extensions that do not exist in the reference are prefixed pnfft_b200_.
 */
#ifndef PNFFT_B200_PNFFT_H
#define PNFFT_B200_PNFFT_H 1

#include <stddef.h>
#include <mpi.h>
#include <pfft.h>   /* the reference's header pulls it in as well (api/pnfft.h:28-29); see include/pfft.h */

#ifdef __cplusplus
extern "C" {
#endif

/* complex numbers follow the FFTW convention the reference inherits through PFFT
 * (reference api/pnfft.h:278-280): C99 complex if <complex.h> was included first, else T[2] */
#if !defined(__cplusplus) && defined(_Complex_I) && defined(complex) && defined(I)
typedef double _Complex pnfft_complex;
typedef float _Complex pnfftf_complex;
#else
typedef double pnfft_complex[2];
typedef float pnfftf_complex[2];
#endif

#define PNFFT_B200_API(PNX, R, C)                                                                   \
  typedef struct PNX(plan_s) *PNX(plan);   /* opaque, reference kernel/ipnfft.h:180-256 */          \
  typedef struct PNX(nodes_s) *PNX(nodes); /* opaque, reference kernel/ipnfft.h:160-177 */          \
                                                                                                    \
  /* process mesh: reference util/util.c:23-40 (0 on success) */                                    \
  int PNX(create_procmesh_2d)(MPI_Comm comm, int np0, int np1, MPI_Comm *comm_cart_2d);             \
  int PNX(create_procmesh)(int rnk, MPI_Comm comm, const int *np, MPI_Comm *comm_cart);             \
                                                                                                    \
  /* block decomposition + node borders: reference api/api-basic.c:36-65, api/api-adv.c:90-140,     \
   * api/api-guru.c:31-107 */                                                                       \
  void PNX(local_size_3d)(const ptrdiff_t *N, MPI_Comm comm_cart, unsigned pnfft_flags,             \
                          ptrdiff_t *local_N, ptrdiff_t *local_N_start, R *lower_border, R *upper_border); \
  void PNX(local_size_3d_c2r)(const ptrdiff_t *N, MPI_Comm comm_cart, unsigned pnfft_flags,         \
                          ptrdiff_t *local_N, ptrdiff_t *local_N_start, R *lower_border, R *upper_border); \
  void PNX(local_size_adv)(int d, const ptrdiff_t *N, MPI_Comm comm_cart, unsigned pnfft_flags,     \
                          ptrdiff_t *local_N, ptrdiff_t *local_N_start, R *lower_border, R *upper_border); \
  void PNX(local_size_adv_c2r)(int d, const ptrdiff_t *N, MPI_Comm comm_cart, unsigned pnfft_flags, \
                          ptrdiff_t *local_N, ptrdiff_t *local_N_start, R *lower_border, R *upper_border); \
  void PNX(local_size_guru)(int d, const ptrdiff_t *N, const ptrdiff_t *n, const R *x_max, int m,   \
                          MPI_Comm comm_cart, unsigned pnfft_flags,                                 \
                          ptrdiff_t *local_N, ptrdiff_t *local_N_start, R *lower_border, R *upper_border); \
  void PNX(local_size_guru_c2r)(int d, const ptrdiff_t *N, const ptrdiff_t *n, const R *x_max, int m, \
                          MPI_Comm comm_cart, unsigned pnfft_flags,                                 \
                          ptrdiff_t *local_N, ptrdiff_t *local_N_start, R *lower_border, R *upper_border); \
                                                                                                    \
  /* plan creation: reference api/api-basic.c:67-98, api/api-adv.c:142-192, api/api-guru.c:64-165 */ \
  PNX(plan) PNX(init_3d)(const ptrdiff_t *N, MPI_Comm comm_cart);                                   \
  PNX(plan) PNX(init_3d_c2r)(const ptrdiff_t *N, MPI_Comm comm_cart);                               \
  PNX(plan) PNX(init_adv)(int d, const ptrdiff_t *N, unsigned pnfft_flags, unsigned fftw_flags, MPI_Comm comm_cart); \
  PNX(plan) PNX(init_adv_c2r)(int d, const ptrdiff_t *N, unsigned pnfft_flags, unsigned fftw_flags, MPI_Comm comm_cart); \
  PNX(plan) PNX(init_guru)(int d, const ptrdiff_t *N, const ptrdiff_t *n, const R *x_max, int m,    \
                           unsigned pnfft_flags, unsigned fftw_flags, MPI_Comm comm_cart);          \
  PNX(plan) PNX(init_guru_c2r)(int d, const ptrdiff_t *N, const ptrdiff_t *n, const R *x_max, int m, \
                           unsigned pnfft_flags, unsigned fftw_flags, MPI_Comm comm_cart);          \
  void PNX(finalize)(PNX(plan) ths, unsigned pnfft_finalize_flags); /* api/api-basic.c:380-400 */   \
                                                                                                    \
  /* node sets: reference api/api-basic.c:402-447 */                                                \
  PNX(nodes) PNX(init_nodes)(ptrdiff_t local_M, unsigned malloc_flags);                             \
  void PNX(free_nodes)(PNX(nodes) ths, unsigned pnfft_finalize_flags);                              \
  /* window precomputation: reference kernel/ndft-parallel.c:1144-1250 */                           \
  void PNX(precompute_psi)(PNX(plan) ths, PNX(nodes) nodes, unsigned precompute_flags);             \
                                                                                                    \
  /* setters / getters: reference api/api-basic.c:457-661 */                                        \
  void PNX(set_f)(C *f, PNX(nodes) nodes);                                                          \
  void PNX(set_grad_f)(C *grad_f, PNX(nodes) nodes);                                                \
  void PNX(set_hessian_f)(C *hessian_f, PNX(nodes) nodes);                                          \
  void PNX(set_f_real)(R *f, PNX(nodes) nodes);                                                     \
  void PNX(set_grad_f_real)(R *grad_f, PNX(nodes) nodes);                                           \
  void PNX(set_hessian_f_real)(R *hessian_f, PNX(nodes) nodes);                                     \
  void PNX(set_x)(R *x, PNX(nodes) nodes);                                                          \
  void PNX(set_f_hat)(C *f_hat, PNX(plan) ths);                                                     \
  void PNX(set_f_hat_real)(R *f_hat, PNX(plan) ths);                                                \
  void PNX(set_b)(R b0, R b1, R b2, PNX(plan) ths);                                                 \
  C *PNX(get_f)(const PNX(nodes) nodes);                                                            \
  C *PNX(get_grad_f)(const PNX(nodes) nodes);                                                       \
  C *PNX(get_hessian_f)(const PNX(nodes) nodes);                                                    \
  R *PNX(get_f_real)(const PNX(nodes) nodes);                                                       \
  R *PNX(get_grad_f_real)(const PNX(nodes) nodes);                                                  \
  R *PNX(get_hessian_f_real)(const PNX(nodes) nodes);                                               \
  R *PNX(get_x)(const PNX(nodes) nodes);                                                            \
  C *PNX(get_f_hat)(const PNX(plan) ths);                                                           \
  R *PNX(get_f_hat_real)(const PNX(plan) ths);                                                      \
  int PNX(get_d)(const PNX(plan) ths);                                                              \
  int PNX(get_m)(const PNX(plan) ths);                                                              \
  void PNX(get_x_max)(const PNX(plan) ths, R *x_max);                                               \
  void PNX(get_N)(const PNX(plan) ths, ptrdiff_t *N);                                               \
  void PNX(get_n)(const PNX(plan) ths, ptrdiff_t *n);                                               \
  unsigned PNX(get_pnfft_flags)(const PNX(plan) ths);                                               \
  unsigned PNX(get_pfft_flags)(const PNX(plan) ths);                                                \
  void PNX(get_b)(const PNX(plan) ths, R *b0, R *b1, R *b2);                                        \
                                                                                                    \
  /* THE HOT PATH: reference api/api-basic.c:199-244 (trafo = B F D) and :344-378 (adj) */          \
  void PNX(trafo)(PNX(plan) ths, PNX(nodes) nodes, unsigned compute_flags);                         \
  void PNX(adj)(PNX(plan) ths, PNX(nodes) nodes, unsigned compute_flags);                           \
                                                                                                    \
  void PNX(init)(void);    /* reference api/api-basic.c:27-30 */                                    \
  void PNX(cleanup)(void); /* reference api/api-basic.c:31-34 */                                    \
  /* page-locked host memory (reference kernel/malloc.c:25-45 forwards to pfft_malloc) */           \
  void *PNX(malloc)(size_t n);                                                                      \
  R *PNX(alloc_real)(size_t n);                                                                     \
  C *PNX(alloc_complex)(size_t n);                                                                  \
  void PNX(free)(void *p);                                                                          \
                                                                                                    \
  /* test-data helpers: reference api/api-basic.c:663-818, api/api-adv.c:35-85 */                   \
  void PNX(init_f_hat_3d)(const ptrdiff_t *N, const ptrdiff_t *local_N, const ptrdiff_t *local_N_start, \
                          unsigned pnfft_flags, C *data);                                           \
  void PNX(init_f)(ptrdiff_t local_M, C *data);                                                     \
  void PNX(init_x_3d)(const R *lo, const R *up, ptrdiff_t loc_M, R *x);                             \
  void PNX(init_x_3d_adv)(const R *lo, const R *up, const R *x_max, ptrdiff_t loc_M, R *x);         \
  void PNX(zero_f_hat)(PNX(plan) ths); /* api/api-basic.c:335-342 */                                \
                                                                                                    \
  /* scalar window functions (host): reference kernel/matrix_D.c:191-225,                           \
   * kernel/ndft-parallel.c:2288-2465 */                                                            \
  R PNX(inv_phi_hat)(const PNX(plan) ths, int dim, ptrdiff_t k);                                    \
  R PNX(phi_hat)(const PNX(plan) ths, int dim, ptrdiff_t k);                                        \
  R PNX(psi)(const PNX(plan) ths, int dim, R x);                                                    \
  R PNX(dpsi)(const PNX(plan) ths, int dim, R x);                                                   \
                                                                                                    \
  void PNX(vpr_complex)(C *data, ptrdiff_t N, const char *name, MPI_Comm comm);                     \
  void PNX(vpr_real)(R *data, ptrdiff_t N, const char *name, MPI_Comm comm);                        \
  /* per-rank print of a 3-d block: reference api/pnfft.h:233-238 */                                \
  void PNX(apr_complex_3d)(C *data, ptrdiff_t *local_N, ptrdiff_t *local_N_start, unsigned pnfft_flags, \
                           const char *name, MPI_Comm comm);                                        \
  void PNX(apr_real_3d)(R *data, ptrdiff_t *local_N, ptrdiff_t *local_N_start, unsigned pnfft_flags, \
                        const char *name, MPI_Comm comm);                                           \
  /* second window derivative (Hessian path, out of scope): prints a notice and returns 0 */         \
  R PNX(ddpsi)(const PNX(plan) ths, int dim, R x);                                                  \
  /* command line helpers of the test drivers: reference util/getargs.c:24-31, api/api-basic.c:820-938 */ \
  void PNX(get_args)(int argc, char **argv, const char *name, int neededArgs, unsigned type, void *parameter); \
  void PNX(check_init_parameters)(int argc, char **argv, ptrdiff_t *N, ptrdiff_t *n, ptrdiff_t *M, int *m, \
                                  unsigned *pnfft_flags, unsigned *compute_flags, double *x_max, int *np, \
                                  int *compare_direct, int *debug);                                 \
                                                                                                    \
  /* timers: reference kernel/timer.c:43-370; the ten slots are filled from CUDA events */          \
  double *PNX(get_timer_trafo)(PNX(plan) ths);                                                      \
  double *PNX(get_timer_adj)(PNX(plan) ths);                                                        \
  void PNX(timer_average)(double *timer);                                                           \
  double *PNX(timer_copy)(const double *orig);                                                      \
  double *PNX(timer_reduce_max)(MPI_Comm comm, double *timer);                                      \
  double *PNX(timer_add)(const double *sum1, const double *sum2);                                   \
  void PNX(timer_free)(double *ths);                                                                \
  void PNX(reset_timer)(PNX(plan) ths);                                                             \
  void PNX(print_average_timer)(const PNX(plan) ths, MPI_Comm comm);                                \
  void PNX(print_average_timer_adv)(const PNX(plan) ths, MPI_Comm comm);                            \
  void PNX(write_average_timer)(const PNX(plan) ths, const char *name, MPI_Comm comm);              \
  void PNX(write_average_timer_adv)(const PNX(plan) ths, const char *name, MPI_Comm comm);          \
                                                                                                    \
  /* ---- extensions (no reference counterpart) -------------------------------------------- */   \
  /* padded-grid access for isolating B (PNFFT_OMIT_DECONV|PNFFT_OMIT_FFT, reference               \
   * api/api-basic.c:176-192): copy the rank's compact local_no block in / out of the device grid */ \
  void PNX(b200_set_grid)(PNX(plan) ths, const R *compact_grid);                                    \
  void PNX(b200_get_grid)(PNX(plan) ths, R *compact_grid);                                          \
  /* compact FFT-input-side array g1 (the rank's local_N block, after D / before D^H) */           \
  void PNX(b200_set_g1)(PNX(plan) ths, const C *g1);                                                \
  void PNX(b200_get_g1)(PNX(plan) ths, C *g1);                                                      \
  void PNX(b200_get_local_no)(const PNX(plan) ths, ptrdiff_t *local_no, ptrdiff_t *local_no_start, ptrdiff_t *no); \
  /* integer parity probes: per node u_j[3] in the padded local array and the plain index m0       \
   * (reference kernel/ndft-parallel.c:1563-1572,2165-2174, ipnfft.h:66), computed on the device */ \
  void PNX(b200_node_grid_index)(PNX(plan) ths, PNX(nodes) nodes, ptrdiff_t *u_and_m0 /* [M][4] */); \
  /* sort key + stable order of reference kernel/ndft-parallel.c:2121-2159, computed on the device */ \
  void PNX(b200_sort_nodes)(PNX(plan) ths, PNX(nodes) nodes, ptrdiff_t *keys /* [M] */, ptrdiff_t *perm /* [M] */); \
  /* 3*(2m+1) window values (and derivatives, may be NULL) per node as the kernels evaluate them */ \
  void PNX(b200_window_tensor)(PNX(plan) ths, PNX(nodes) nodes, R *psi /* [M][3][2m+1] */, R *dpsi); \
  /* select kernels (default 0 = z-marching register kernels): bit 0 = generic global-memory gridding  \
   * kernels; bit 1 = exact window evaluation instead of the per-tap polynomials fitted at plan time; \
   * bit 3 = z-marching v1 (CTA-synchronous) kernels instead of the warp-autonomous v2 */            \
  void PNX(b200_set_kernel_variant)(PNX(plan) ths, int variant);                                    \
  /* promise that the node coordinates stay unchanged until the next pnfft_set_x: their upload and   \
   * binning are then done once and reused by every pnfft_trafo / pnfft_adj on these nodes (the      \
   * reference re-reads x every call, api/api-basic.c:199-244).  PNFFT_B200_X_STATIC=1 sets it for    \
   * all node sets of an unmodified caller */                                                         \
  void PNX(b200_nodes_x_static)(PNX(nodes) nodes, int on);                                           \
  int PNX(b200_get_poly_degree)(PNX(plan) ths);                                                     \
  /* device time (ms) of the last trafo/adj stages: [0]=B gather/scatter kernel only,              \
   * [1]=binning, [2]=halo, [3]=F, [4]=D, [5]=H2D, [6]=D2H, [7]=whole */                            \
  void PNX(b200_get_stage_ms)(PNX(plan) ths, int adjoint, double *ms8);                             \
  /* kernels of this library launched so far / cuFFT, CUB and NCCL calls issued so far */        \
  long long PNX(b200_kernel_launches)(PNX(plan) ths);                                               \
  long long PNX(b200_library_calls)(PNX(plan) ths);                                                 \
  /* host evaluation of the Kaiser-Bessel taps as the double-precision node-table kernel computes them:      \
   * psi / dpsi are [M][3][2m+1] for nodes x[M][3], oversampled sizes n[3], shape parameters b[3] */           \
  void PNX(b200_kb_taps_host)(const double *x, ptrdiff_t M, const ptrdiff_t *n, const double *b, int m, double *psi, double *dpsi); \
  /* host evaluation of the window's Fourier coefficients as the D tables hold them: out[i] = phi_hat(k[i]) or, with       \
   * inverse != 0, 1 / phi_hat(k[i]) for the window of pnfft_flags (b <= 0: that window's default shape at sigma = n / N) */ \
  void PNX(b200_phi_hat_host)(unsigned pnfft_flags, ptrdiff_t N, ptrdiff_t n, R b, int m, const ptrdiff_t *k, ptrdiff_t len, int inverse, R *out); \
  /* host evaluation of pnfft_psi (which = 0), pnfft_dpsi (1), pnfft_ddpsi (2) at offsets x[len] for the window of pnfft_flags */ \
  void PNX(b200_psi_host)(unsigned pnfft_flags, ptrdiff_t N, ptrdiff_t n, R b, int m, int which, const R *x, ptrdiff_t len, R *out); \
  /* host-only: the f_hat block of rank pid of a p0 x p1 mesh as the direct NDFT exchanges it, memory order:            \
   * out12 = { len[3], start[3], axis[3], N_of_axis[3] } */                                                             \
  void PNX(b200_direct_block)(const ptrdiff_t *N, const ptrdiff_t *n, int m, int p0, int p1, int pid, unsigned pnfft_flags, int c2r, int *out12); \
  /* host-only: append the report of pnfft_write_average_timer (adv = 0) / _adv (adv != 0) for the given plan data to `name` */ \
  void PNX(b200_timer_report_host)(const char *name, unsigned pnfft_flags, const ptrdiff_t *N, const ptrdiff_t *n, int m,      \
                                   const int *np3, const double *timer_trafo, const double *timer_adj, int adv, MPI_Comm comm); \
  /* host-only self check of the pencil FFT's composed "own chunk" maps for rank (c0, c1) of a p0 x p1 \
   * mesh: self transfers checked, -1 on a mismatch, -2 if one did not compose (no GPU needed) */       \
  int PNX(b200_check_self_maps)(const ptrdiff_t *N, const ptrdiff_t *n, int m, int p0, int p1, int c0, int c1, int c2r); \
  /* host-only: the z range [tz0, tz1) of work item `seg` of a column whose bins (nt2 sub-chunks x sub x-offset bins, prefix \
   * sums `prefix[nt2 * sub + 1]`) are cut into up to nseg pieces as the gridding kernels do (target <= 0: equal length) */  \
  void PNX(b200_column_piece)(const int *prefix, int nt2, int sub, int seg, int nseg, int target, int fill, int *tz0, int *tz1); \
  /* the plan's cudaStream_t (all work of trafo/adj is issued on it; calls synchronise it on return) */ \
  void *PNX(b200_get_stream)(PNX(plan) ths);

#define PNFFT_B200_MANGLE_D(name) pnfft_##name
#define PNFFT_B200_MANGLE_F(name) pnfftf_##name

PNFFT_B200_API(PNFFT_B200_MANGLE_D, double, pnfft_complex)
PNFFT_B200_API(PNFFT_B200_MANGLE_F, float, pnfftf_complex)

/* FP64 FMA rate (TFLOP/s) of the current device, measured with independent DFMA chains on all SMs:
 * the compute denominator of the gridding roofline (bench.py) */
double pnfft_b200_measure_fp64_tflops(void);

#ifndef PNFFT_PI
#define PNFFT_PI 3.14159265358979323846
#endif

/* ---- plan flags (values are ABI: reference api/pnfft.h:302-334) ---- */
#define PNFFT_PRE_PHI_HAT          (1U << 0)
#define PNFFT_PRE_PHI_HUT          (PNFFT_PRE_PHI_HAT)
#define PNFFT_FAST_GAUSSIAN        (1U << 1)
#define PNFFT_FG_PSI               (PNFFT_FAST_GAUSSIAN) /* name used by the reference manual, doc/manual.tex:268 */
#define PNFFT_PRE_CONST_PSI        (1U << 2)
#define PNFFT_PRE_LIN_PSI          (1U << 3)
#define PNFFT_PRE_QUAD_PSI         (1U << 4)
#define PNFFT_PRE_CUB_PSI          (1U << 5)
#define PNFFT_PRE_INTPOL_PSI       (PNFFT_PRE_CONST_PSI | PNFFT_PRE_LIN_PSI | PNFFT_PRE_QUAD_PSI | PNFFT_PRE_CUB_PSI)
#define PNFFT_MALLOC_F_HAT         (1U << 6)
#define PNFFT_FFT_OUT_OF_PLACE     (0U)
#define PNFFT_FFT_IN_PLACE         (1U << 7)
#define PNFFT_INTERLACED           (1U << 8)
#define PNFFT_SHIFTED_F_HAT        (1U << 9)
#define PNFFT_SHIFTED_X            (1U << 10)
#define PNFFT_TRANSPOSED_NONE      (0U)
#define PNFFT_TRANSPOSED_F_HAT     (1U << 11)
#define PNFFT_DIFF_AD              (0U)
#define PNFFT_DIFF_IK              (1U << 12)
#define PNFFT_WINDOW_KAISER_BESSEL (0U)
#define PNFFT_WINDOW_GAUSSIAN      (1U << 13)
#define PNFFT_WINDOW_BSPLINE       (1U << 14)
#define PNFFT_WINDOW_SINC_POWER    (1U << 15)
#define PNFFT_WINDOW_BESSEL_I0     (1U << 16)
#define PNFFT_USE_FK_GAUSSIAN_T    (1U << 17)
#define PNFFT_WINDOW_GAUSSIAN_T    (PNFFT_USE_FK_GAUSSIAN_T | PNFFT_WINDOW_GAUSSIAN)
#define PNFFT_SORT_NODES           (1U << 18)

/* ---- finalize / node flags (reference api/pnfft.h:340-377) ---- */
#define PNFFT_FREE_F_HAT       (PNFFT_MALLOC_F_HAT)
#define PNFFT_MALLOC_NONE      (0U)
#define PNFFT_MALLOC_X         (1U << 0)
#define PNFFT_MALLOC_F         (1U << 1)
#define PNFFT_MALLOC_GRAD_F    (1U << 2)
#define PNFFT_MALLOC_HESSIAN_F (1U << 3)
#define PNFFT_MALLOC_ALL       (PNFFT_MALLOC_X | PNFFT_MALLOC_F | PNFFT_MALLOC_GRAD_F | PNFFT_MALLOC_HESSIAN_F)
#define PNFFT_REAL_F           (1U << 4)
#define PNFFT_FREE_NONE        (0U)
#define PNFFT_FREE_X           (PNFFT_MALLOC_X)
#define PNFFT_FREE_F           (PNFFT_MALLOC_F)
#define PNFFT_FREE_GRAD_F      (PNFFT_MALLOC_GRAD_F)
#define PNFFT_FREE_HESSIAN_F   (PNFFT_MALLOC_HESSIAN_F)
#define PNFFT_FREE_ALL         (PNFFT_MALLOC_ALL)

/* ---- precompute flags (reference api/pnfft.h:359-364) ---- */
#define PNFFT_PRE_TENSOR       (0U)
#define PNFFT_PRE_FULL         (1U << 0)
#define PNFFT_PRE_PSI          (1U << 1)
#define PNFFT_PRE_GRAD_PSI     (1U << 2)
#define PNFFT_PRE_HESSIAN_PSI  (1U << 3)

/* ---- compute flags (reference api/pnfft.h:383-390) ---- */
#define PNFFT_COMPUTE_F           (1U << 0)
#define PNFFT_COMPUTE_GRAD_F      (1U << 1)
#define PNFFT_COMPUTE_HESSIAN_F   (1U << 2)
#define PNFFT_COMPUTE_DIRECT      (1U << 3)
#define PNFFT_COMPUTE_ACCUMULATED (1U << 4)
#define PNFFT_OMIT_DECONV         (1U << 5)
#define PNFFT_OMIT_FFT            (1U << 6)
#define PNFFT_OMIT_CONV           (1U << 7)

/* ---- timer slots (reference api/pnfft.h:407-418) ---- */
#define PNFFT_TIMER_ITER         (0)
#define PNFFT_TIMER_WHOLE        (1)
#define PNFFT_TIMER_LOOP_B       (2)
#define PNFFT_TIMER_SORT_NODES   (3)
#define PNFFT_TIMER_GCELLS       (4)
#define PNFFT_TIMER_MATRIX_B     (5)
#define PNFFT_TIMER_MATRIX_F     (6)
#define PNFFT_TIMER_MATRIX_D     (7)
#define PNFFT_TIMER_SHIFT_INPUT  (8)
#define PNFFT_TIMER_SHIFT_OUTPUT (9)
#define PNFFT_TIMER_LENGTH       (10)

#ifdef __cplusplus
}
#endif
#endif
