/* pnfft-b200: the part of <pfft.h> that PNFFT callers and the reference's test drivers touch directly
 * (pfft_printf, pfft_fprintf, pfft_get_args, pfft_prod_INT, pfft_apr_complex_3d, pfft_complex and the
 * PFFT_* constants handed to pnfft_init_guru / pfft_get_args).  PFFT itself is not needed: block decomposition, the
 * pruned FFT and the ghost cells live inside libpnfft_b200.so.  The planner flags are accepted and ignored. */
#ifndef PNFFT_B200_PFFT_H
#define PNFFT_B200_PFFT_H 1

#include <stddef.h>
#include <stdio.h>
#include <mpi.h>

#ifdef __cplusplus
extern "C" {
#endif

#if !defined(__cplusplus) && defined(_Complex_I) && defined(complex) && defined(I)
typedef double _Complex pfft_complex;
typedef float _Complex pfftf_complex;
#else
typedef double pfft_complex[2];
typedef float pfftf_complex[2];
#endif

/* argument types of pfft_get_args */
#define PFFT_INT         (1U)
#define PFFT_PTRDIFF_T   (2U)
#define PFFT_FLOAT       (3U)
#define PFFT_DOUBLE      (4U)
#define PFFT_UNSIGNED    (5U)

/* planner / layout flags (values are private to this header: they only travel through the pfft_flags argument) */
#define PFFT_MEASURE            (0U)
#define PFFT_ESTIMATE           (1U << 2)
#define PFFT_PATIENT            (1U << 3)
#define PFFT_EXHAUSTIVE         (1U << 4)
#define PFFT_DESTROY_INPUT      (1U << 5)
#define PFFT_PRESERVE_INPUT     (1U << 6)
#define PFFT_TRANSPOSED_NONE    (0U)
#define PFFT_TRANSPOSED_IN      (1U << 7)
#define PFFT_TRANSPOSED_OUT     (1U << 8)
#define PFFT_SHIFTED_NONE       (0U)
#define PFFT_SHIFTED_IN         (1U << 9)
#define PFFT_SHIFTED_OUT        (1U << 10)

/* rank 0 of comm prints (printf semantics) */
void pfft_printf(MPI_Comm comm, const char *format, ...);
void pfft_fprintf(MPI_Comm comm, FILE *stream, const char *format, ...);
/* command line: "-name v1 .. vn" fills n values of the given type; missing option leaves the defaults untouched */
void pfft_get_args(int argc, char **argv, const char *name, const int neededArgs, const unsigned type, void *parameter);
ptrdiff_t pfft_prod_INT(int d, const ptrdiff_t *vec);
/* every rank in turn prints its block of a 3-d complex array */
void pfft_apr_complex_3d(const pfft_complex *data, const ptrdiff_t *local_n, const ptrdiff_t *local_start, const char *name,
                         MPI_Comm comm);

/* the float instantiation of PFFT prefixes the same helpers with pfftf_ (reference tests/check_trafo_vs_ndft_float.c) */
void pfftf_printf(MPI_Comm comm, const char *format, ...);
void pfftf_fprintf(MPI_Comm comm, FILE *stream, const char *format, ...);
void pfftf_get_args(int argc, char **argv, const char *name, const int neededArgs, const unsigned type, void *parameter);
ptrdiff_t pfftf_prod_INT(int d, const ptrdiff_t *vec);

#ifdef __cplusplus
}
#endif
#endif
