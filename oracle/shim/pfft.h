/* TEST INFRASTRUCTURE ONLY -- part of the oracle, never linked into the product.
 *
 * Stand-in for PFFT's <pfft.h> (PFFT >= 1.0.8-alpha is an external dependency of
 * the reference and is NOT in /root/reference).  Declares exactly the 37 pfft_*
 * symbols the reference's window-convolution path references plus
 * what its test drivers use.  The semantics implemented in shim_pfft.c restate
 * PFFT's documented behaviour (doc/manual.tex:178-238 of the reference, PFFT user
 * manual): default block decomposition ceil(n/P), shifted index ranges, pruned
 * input/output, periodic ghost cells.  Exact PFFT output for uneven splits is
 * "parity unpinned" (no PFFT source, no reference test asserts it).
 */
#ifndef ORACLE_SHIM_PFFT_H
#define ORACLE_SHIM_PFFT_H 1

#include <stddef.h>
#include <stdio.h>
#include <math.h>
#include <mpi.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFFT_CONCAT(prefix, name) prefix ## name
#define PFFT_MANGLE_DOUBLE(name) PFFT_CONCAT(pfft_, name)
#define PFFT_MANGLE_FLOAT(name) PFFT_CONCAT(pfftf_, name)
#define PFFT_MANGLE_LONG_DOUBLE(name) PFFT_CONCAT(pfftl_, name)
#define FFTW_MANGLE_DOUBLE(name) PFFT_CONCAT(fftw_, name)
#define FFTW_MANGLE_FLOAT(name) PFFT_CONCAT(fftwf_, name)
#define FFTW_MANGLE_LONG_DOUBLE(name) PFFT_CONCAT(fftwl_, name)

typedef double _Complex pfft_complex;
typedef float _Complex pfftf_complex;
typedef long double _Complex pfftl_complex;

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define PFFT_FORWARD  (FFTW_FORWARD)
#define PFFT_BACKWARD (FFTW_BACKWARD)

#define PFFT_DEFAULT_BLOCKS  NULL
#define PFFT_DEFAULT_BLOCK   ((ptrdiff_t)-1)

/* plan flags (values are private to the shim; the reference only uses the names) */
#define PFFT_TRANSPOSED_NONE (0U)
#define PFFT_TRANSPOSED_IN   (1U << 0)
#define PFFT_TRANSPOSED_OUT  (1U << 1)
#define PFFT_SHIFTED_NONE    (0U)
#define PFFT_SHIFTED_IN      (1U << 2)
#define PFFT_SHIFTED_OUT     (1U << 3)
#define PFFT_MEASURE         (0U)
#define PFFT_ESTIMATE        (1U << 4)
#define PFFT_PATIENT         (1U << 5)
#define PFFT_EXHAUSTIVE      (1U << 6)
#define PFFT_NO_TUNE         (0U)
#define PFFT_TUNE            (1U << 7)
#define PFFT_PRESERVE_INPUT  (1U << 8)
#define PFFT_DESTROY_INPUT   (1U << 9)
#define PFFT_BUFFERED_INPLACE (1U << 10)
#define PFFT_PADDED_R2C      (1U << 11)
#define PFFT_PADDED_C2R      (1U << 12)

/* types for pfft_get_args */
#define PFFT_INT       (1U)
#define PFFT_PTRDIFF_T (2U)
#define PFFT_FLOAT     (3U)
#define PFFT_DOUBLE    (4U)
#define PFFT_UNSIGNED  (5U)
#define PFFT_LDOUBLE   (6U)

#define PFFT_SHIM_DEFINE_API(PX, R, C, INT)                                              \
  typedef struct PX(plan_s) *PX(plan);                                                   \
  typedef struct PX(gcplan_s) *PX(gcplan);                                               \
  void PX(init)(void);                                                                   \
  void PX(cleanup)(void);                                                                \
  void *PX(malloc)(size_t n);                                                            \
  R *PX(alloc_real)(size_t n);                                                           \
  C *PX(alloc_complex)(size_t n);                                                        \
  void PX(free)(void *p);                                                                \
  int PX(create_procmesh)(int rnk, MPI_Comm comm, const int *np, MPI_Comm *comm_cart);   \
  int PX(create_procmesh_2d)(MPI_Comm comm, int np0, int np1, MPI_Comm *comm_cart_2d);   \
  INT PX(local_size_many_dft)(int rnk_n, const INT *n, const INT *ni, const INT *no,     \
      INT howmany, const INT *iblock, const INT *oblock, MPI_Comm comm_cart,             \
      unsigned pfft_flags, INT *local_ni, INT *local_i_start, INT *local_no,             \
      INT *local_o_start);                                                               \
  INT PX(local_size_many_dft_c2r)(int rnk_n, const INT *n, const INT *ni, const INT *no, \
      INT howmany, const INT *iblock, const INT *oblock, MPI_Comm comm_cart,             \
      unsigned pfft_flags, INT *local_ni, INT *local_i_start, INT *local_no,             \
      INT *local_o_start);                                                               \
  INT PX(local_size_many_dft_r2c)(int rnk_n, const INT *n, const INT *ni, const INT *no, \
      INT howmany, const INT *iblock, const INT *oblock, MPI_Comm comm_cart,             \
      unsigned pfft_flags, INT *local_ni, INT *local_i_start, INT *local_no,             \
      INT *local_o_start);                                                               \
  void PX(local_block_many_dft)(int rnk_n, const INT *ni, const INT *no,                 \
      const INT *iblock, const INT *oblock, MPI_Comm comm_cart, int pid,                 \
      unsigned pfft_flags, INT *local_ni, INT *local_i_start, INT *local_no,             \
      INT *local_o_start);                                                               \
  void PX(local_block_many_dft_c2r)(int rnk_n, const INT *ni, const INT *no,             \
      const INT *iblock, const INT *oblock, MPI_Comm comm_cart, int pid,                 \
      unsigned pfft_flags, INT *local_ni, INT *local_i_start, INT *local_no,             \
      INT *local_o_start);                                                               \
  INT PX(local_size_many_gc)(int rnk_n, const INT *local_n, const INT *local_n_start,    \
      INT howmany, const INT *gc_below, const INT *gc_above, INT *local_ngc,             \
      INT *local_gc_start);                                                              \
  PX(plan) PX(plan_many_dft)(int rnk_n, const INT *n, const INT *ni, const INT *no,      \
      INT howmany, const INT *iblock, const INT *oblock, C *in, C *out,                  \
      MPI_Comm comm_cart, int sign, unsigned pfft_flags);                                \
  PX(plan) PX(plan_many_dft_c2r)(int rnk_n, const INT *n, const INT *ni, const INT *no,  \
      INT howmany, const INT *iblock, const INT *oblock, C *in, R *out,                  \
      MPI_Comm comm_cart, int sign, unsigned pfft_flags);                                \
  PX(plan) PX(plan_many_dft_r2c)(int rnk_n, const INT *n, const INT *ni, const INT *no,  \
      INT howmany, const INT *iblock, const INT *oblock, R *in, C *out,                  \
      MPI_Comm comm_cart, int sign, unsigned pfft_flags);                                \
  PX(gcplan) PX(plan_many_cgc)(int rnk_n, const INT *n, INT howmany, const INT *block,   \
      const INT *gc_below, const INT *gc_above, C *data, MPI_Comm comm_cart,             \
      unsigned gc_flags);                                                                \
  PX(gcplan) PX(plan_many_rgc)(int rnk_n, const INT *n, INT howmany, const INT *block,   \
      const INT *gc_below, const INT *gc_above, R *data, MPI_Comm comm_cart,             \
      unsigned gc_flags);                                                                \
  void PX(execute)(const PX(plan) ths);                                                  \
  void PX(exchange)(PX(gcplan) ths);                                                     \
  void PX(reduce)(PX(gcplan) ths);                                                       \
  void PX(destroy_plan)(PX(plan) ths);                                                   \
  void PX(destroy_gcplan)(PX(gcplan) ths);                                               \
  INT PX(prod_INT)(int d, const INT *vec);                                               \
  INT PX(sum_INT)(int d, const INT *vec);                                                \
  int PX(equal_INT)(int d, const INT *vec1, const INT *vec2);                            \
  void PX(vcopy_INT)(int d, const INT *vec1, INT *vec2);                                 \
  void PX(vadd_INT)(int d, const INT *vec1, const INT *vec2, INT *sum);                  \
  void PX(vsub_INT)(int d, const INT *vec1, const INT *vec2, INT *sum);                  \
  void PX(fprintf)(MPI_Comm comm, FILE *stream, const char *format, ...);                \
  void PX(printf)(MPI_Comm comm, const char *format, ...);                               \
  void PX(get_args)(int argc, char **argv, const char *name, int neededArgs,             \
      unsigned type, void *parameter);                                                   \
  void PX(init_input_complex_3d)(const INT *n, const INT *local_n,                       \
      const INT *local_n_start, C *data);                                                \
  void PX(apr_complex_3d)(const C *data, const INT *local_n, const INT *local_n_start,   \
      const char *name, MPI_Comm comm);                                                  \
  void PX(apr_real_3d)(const R *data, const INT *local_n, const INT *local_n_start,      \
      const char *name, MPI_Comm comm);                                                  \
  void PX(print_average_timer_adv)(const PX(plan) ths, MPI_Comm comm);                   \
  void PX(write_average_timer_adv)(const PX(plan) ths, const char *name, MPI_Comm comm); \
  void PX(print_average_gctimer_adv)(const PX(gcplan) ths, MPI_Comm comm);               \
  void PX(write_average_gctimer_adv)(const PX(gcplan) ths, const char *name, MPI_Comm comm);

PFFT_SHIM_DEFINE_API(PFFT_MANGLE_DOUBLE, double, pfft_complex, ptrdiff_t)
PFFT_SHIM_DEFINE_API(PFFT_MANGLE_FLOAT, float, pfftf_complex, ptrdiff_t)

#ifdef __cplusplus
}
#endif
#endif
