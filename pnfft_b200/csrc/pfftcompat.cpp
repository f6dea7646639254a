// The handful of PFFT helper entry points that PNFFT callers and the reference's test drivers call directly
// (include/pfft.h).  Behaviour restated from their use in reference tests/*.c and the PFFT manual:
// rank-0 printing, "-name v1 .. vn" command-line parsing, integer vector product, per-rank array print.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <pfft.h>

extern "C" {

static int rank_of(MPI_Comm comm) {
  int r = 0;
  MPI_Comm_rank(comm, &r);
  return r;
}

void pfft_printf(MPI_Comm comm, const char *format, ...) {
  if (rank_of(comm) != 0) return;
  va_list ap;
  va_start(ap, format);
  vfprintf(stdout, format, ap);
  va_end(ap);
  fflush(stdout);
}

void pfft_fprintf(MPI_Comm comm, FILE *stream, const char *format, ...) {
  if (rank_of(comm) != 0) return;
  va_list ap;
  va_start(ap, format);
  vfprintf(stream, format, ap);
  va_end(ap);
  fflush(stream);
}

void pfft_get_args(int argc, char **argv, const char *name, const int neededArgs, const unsigned type, void *parameter) {
  for (int i = 1; i < argc; i++) {
    if (strcmp(argv[i], name) != 0) continue;
    if (i + neededArgs > argc - 1) return;   // not enough values: keep the defaults
    for (int k = 0; k < neededArgs; k++) {
      const char *v = argv[i + 1 + k];
      switch (type) {
        case PFFT_INT: ((int *)parameter)[k] = atoi(v); break;
        case PFFT_PTRDIFF_T: ((ptrdiff_t *)parameter)[k] = (ptrdiff_t)atoll(v); break;
        case PFFT_FLOAT: ((float *)parameter)[k] = (float)atof(v); break;
        case PFFT_DOUBLE: ((double *)parameter)[k] = atof(v); break;
        case PFFT_UNSIGNED: ((unsigned *)parameter)[k] = (unsigned)strtoul(v, nullptr, 10); break;
        default: break;
      }
    }
    return;
  }
}

void pfftf_printf(MPI_Comm comm, const char *format, ...) {
  if (rank_of(comm) != 0) return;
  va_list ap;
  va_start(ap, format);
  vfprintf(stdout, format, ap);
  va_end(ap);
  fflush(stdout);
}
void pfftf_fprintf(MPI_Comm comm, FILE *stream, const char *format, ...) {
  if (rank_of(comm) != 0) return;
  va_list ap;
  va_start(ap, format);
  vfprintf(stream, format, ap);
  va_end(ap);
  fflush(stream);
}
void pfftf_get_args(int argc, char **argv, const char *name, const int neededArgs, const unsigned type, void *parameter) {
  pfft_get_args(argc, argv, name, neededArgs, type, parameter);
}
ptrdiff_t pfftf_prod_INT(int d, const ptrdiff_t *vec) { return pfft_prod_INT(d, vec); }

ptrdiff_t pfft_prod_INT(int d, const ptrdiff_t *vec) {
  ptrdiff_t p = 1;
  for (int t = 0; t < d; t++) p *= vec[t];
  return p;
}

void pfft_apr_complex_3d(const pfft_complex *data, const ptrdiff_t *local_n, const ptrdiff_t *local_start, const char *name,
                         MPI_Comm comm) {
  int rank = 0, size = 1;
  MPI_Comm_rank(comm, &rank);
  MPI_Comm_size(comm, &size);
  const double *d = (const double *)data;
  for (int r = 0; r < size; r++) {
    if (r == rank) {
      printf("rank %d: %s", rank, name);
      ptrdiff_t l = 0;
      for (ptrdiff_t k0 = 0; k0 < local_n[0]; k0++)
        for (ptrdiff_t k1 = 0; k1 < local_n[1]; k1++) {
          for (ptrdiff_t k2 = 0; k2 < local_n[2]; k2++, l++)
            printf("  [%td,%td,%td] %.4e%+.4ei", k0 + local_start[0], k1 + local_start[1], k2 + local_start[2], d[2 * l], d[2 * l + 1]);
          printf("\n");
        }
      fflush(stdout);
    }
    MPI_Barrier(comm);
  }
}

}  // extern "C"
