/* TEST INFRASTRUCTURE ONLY -- CPU emulation of the PFFT calls the reference PNFFT
 * makes (PFFT is an external dependency that is not in /root/reference).
 *
 * Compiled twice (double: default, float: -DPNFFT_PREC_SINGLE) into
 * oracle/_ref/libpnfft_ref.so.  Semantics restated from PFFT's documented
 * behaviour (reference doc/manual.tex:178-238, doc/intro.tex:32-58):
 *   - block decomposition over a 1-d/2-d process mesh with default block
 *     ceil(n/P); rank c owns [c*block, min(n,(c+1)*block));
 *   - PFFT_SHIFTED_IN/OUT: index ranges [-n/2, n/2) instead of [0, n);
 *   - pruned transforms: only ni inputs / no outputs of a length-n DFT;
 *   - PFFT_TRANSPOSED_IN/OUT: memory order n1 x n2 x n0, split over (n1,n2);
 *   - ghost cells: periodic halo in all three dims of the no-array.
 * All "ranks" are threads of one process (shim_mpi.c), so collectives are done by
 * assembling global arrays in shared memory.  The FFT itself is a plain
 * mixed-radix host FFT evaluated in double precision for both builds.
 */
#include <complex.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "pfft.h"

typedef ptrdiff_t INT;

#if defined(PNFFT_PREC_SINGLE)
typedef float R;
typedef pfftf_complex C;
#define PX(name) PFFT_MANGLE_FLOAT(name)
#else
typedef double R;
typedef pfft_complex C;
#define PX(name) PFFT_MANGLE_DOUBLE(name)
#endif

typedef double _Complex Z;

/* ------------------------------------------------------------------ */
/* small helpers                                                       */
/* ------------------------------------------------------------------ */
void PX(init)(void) {}
void PX(cleanup)(void) {}

void *PX(malloc)(size_t n)
{
  void *p = NULL;
  if (n == 0) n = 1;
  if (posix_memalign(&p, 64, n)) return NULL;
  return p;
}
R *PX(alloc_real)(size_t n) { return (R *)PX(malloc)(sizeof(R) * n); }
C *PX(alloc_complex)(size_t n) { return (C *)PX(malloc)(sizeof(C) * n); }
void PX(free)(void *p) { free(p); }

INT PX(prod_INT)(int d, const INT *v) { INT p = 1; for (int t = 0; t < d; t++) p *= v[t]; return p; }
INT PX(sum_INT)(int d, const INT *v) { INT s = 0; for (int t = 0; t < d; t++) s += v[t]; return s; }
int PX(equal_INT)(int d, const INT *a, const INT *b) { for (int t = 0; t < d; t++) if (a[t] != b[t]) return 0; return 1; }
void PX(vcopy_INT)(int d, const INT *a, INT *b) { for (int t = 0; t < d; t++) b[t] = a[t]; }
void PX(vadd_INT)(int d, const INT *a, const INT *b, INT *s) { for (int t = 0; t < d; t++) s[t] = a[t] + b[t]; }
void PX(vsub_INT)(int d, const INT *a, const INT *b, INT *s) { for (int t = 0; t < d; t++) s[t] = a[t] - b[t]; }

void PX(fprintf)(MPI_Comm comm, FILE *stream, const char *format, ...)
{
  int rank; MPI_Comm_rank(comm, &rank);
  if (rank != 0) return;
  va_list ap; va_start(ap, format); vfprintf(stream, format, ap); va_end(ap);
}
void PX(printf)(MPI_Comm comm, const char *format, ...)
{
  int rank; MPI_Comm_rank(comm, &rank);
  if (rank != 0) return;
  va_list ap; va_start(ap, format); vfprintf(stdout, format, ap); va_end(ap);
}

void PX(get_args)(int argc, char **argv, const char *name, int needed, unsigned type, void *parameter)
{
  for (int i = 1; i < argc; i++) {
    if (strcmp(argv[i], name) != 0) continue;
    if (i + needed > argc - 1) return;
    for (int a = 0; a < needed; a++) {
      const char *s = argv[i + 1 + a];
      switch (type) {
        case PFFT_INT: ((int *)parameter)[a] = atoi(s); break;
        case PFFT_PTRDIFF_T: ((ptrdiff_t *)parameter)[a] = (ptrdiff_t)atoll(s); break;
        case PFFT_FLOAT: ((float *)parameter)[a] = (float)atof(s); break;
        case PFFT_DOUBLE: ((double *)parameter)[a] = atof(s); break;
        case PFFT_UNSIGNED: ((unsigned *)parameter)[a] = (unsigned)strtoul(s, NULL, 10); break;
        case PFFT_LDOUBLE: ((long double *)parameter)[a] = strtold(s, NULL); break;
      }
    }
    return;
  }
}

int PX(create_procmesh)(int rnk, MPI_Comm comm, const int *np, MPI_Comm *comm_cart)
{
  int periods[3] = {1, 1, 1};
  int size, prod = 1;
  MPI_Comm_size(comm, &size);
  for (int t = 0; t < rnk; t++) prod *= np[t];
  if (prod != size) return 1;
  return MPI_Cart_create(comm, rnk, np, periods, 1, comm_cart);
}

int PX(create_procmesh_2d)(MPI_Comm comm, int np0, int np1, MPI_Comm *comm_cart_2d)
{
  int np[2] = {np0, np1};
  return PX(create_procmesh)(2, comm, np, comm_cart_2d);
}

/* ------------------------------------------------------------------ */
/* block decomposition                                                 */
/* ------------------------------------------------------------------ */
static void block_1d(INT n, int p, int c, INT *len, INT *start)
{
  INT blk = (n + p - 1) / p;
  INT s = (INT)c * blk;
  INT l = n - s;
  if (l > blk) l = blk;
  if (l <= 0) { l = 0; s = 0; }
  *len = l; *start = s;
}

/* Decompose a logical n0 x n1 x n2 array for the rank with mesh coordinates co[]
 * of a mesh np[] (np[2] must be 1).  half_last: the array stores only n2/2+1
 * entries of the last dimension (complex side of r2c/c2r). */
static void decompose(const INT *n, const int *np, const int *co, int transposed, int shifted,
                      int half_last, INT *ln, INT *ls)
{
  INT ext[3] = {n[0], n[1], half_last ? n[2] / 2 + 1 : n[2]};
  if (np[2] != 1) { fprintf(stderr, "shim PFFT: 3-d process meshes are not emulated\n"); abort(); }
  if (!transposed) {
    block_1d(ext[0], np[0], co[0], &ln[0], &ls[0]);
    block_1d(ext[1], np[1], co[1], &ln[1], &ls[1]);
    ln[2] = ext[2]; ls[2] = 0;
  } else {
    block_1d(ext[1], np[0], co[0], &ln[1], &ls[1]);
    block_1d(ext[2], np[1], co[1], &ln[2], &ls[2]);
    ln[0] = ext[0]; ls[0] = 0;
  }
  if (shifted) for (int t = 0; t < 3; t++) ls[t] -= n[t] / 2;
}

static void mesh_of(MPI_Comm comm, int rank, int *np, int *co)
{
  int nd, dims[3], me[3];
  shim_comm_dims(comm, &nd, dims, me);
  for (int t = 0; t < 3; t++) np[t] = dims[t];
  if (rank < 0) for (int t = 0; t < 3; t++) co[t] = me[t];
  else shim_rank_coords(comm, rank, co);
}

static INT local_size_generic(const INT *ni, const INT *no, MPI_Comm comm, int pid, unsigned flags,
                              int half_in, int half_out,
                              INT *lni, INT *lis, INT *lno, INT *los)
{
  int np[3], co[3];
  mesh_of(comm, pid, np, co);
  decompose(ni, np, co, (flags & PFFT_TRANSPOSED_IN) != 0, (flags & PFFT_SHIFTED_IN) != 0, half_in, lni, lis);
  decompose(no, np, co, (flags & PFFT_TRANSPOSED_OUT) != 0, (flags & PFFT_SHIFTED_OUT) != 0, half_out, lno, los);
  INT a = lni[0] * lni[1] * lni[2], b = lno[0] * lno[1] * lno[2];
  /* units of complex: a real array of b entries needs (b+1)/2 complex */
  if (half_in && !half_out) b = (b + 1) / 2;
  if (half_out && !half_in) a = (a + 1) / 2;
  return a > b ? a : b;
}

INT PX(local_size_many_dft)(int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
    const INT *iblock, const INT *oblock, MPI_Comm comm, unsigned flags,
    INT *lni, INT *lis, INT *lno, INT *los)
{
  (void)rnk_n; (void)n; (void)iblock; (void)oblock;
  return howmany * local_size_generic(ni, no, comm, -1, flags, 0, 0, lni, lis, lno, los);
}

INT PX(local_size_many_dft_c2r)(int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
    const INT *iblock, const INT *oblock, MPI_Comm comm, unsigned flags,
    INT *lni, INT *lis, INT *lno, INT *los)
{
  (void)rnk_n; (void)n; (void)iblock; (void)oblock;
  return howmany * local_size_generic(ni, no, comm, -1, flags, 1, 0, lni, lis, lno, los);
}

INT PX(local_size_many_dft_r2c)(int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
    const INT *iblock, const INT *oblock, MPI_Comm comm, unsigned flags,
    INT *lni, INT *lis, INT *lno, INT *los)
{
  (void)rnk_n; (void)n; (void)iblock; (void)oblock;
  return howmany * local_size_generic(ni, no, comm, -1, flags, 0, 1, lni, lis, lno, los);
}

void PX(local_block_many_dft)(int rnk_n, const INT *ni, const INT *no, const INT *iblock,
    const INT *oblock, MPI_Comm comm, int pid, unsigned flags,
    INT *lni, INT *lis, INT *lno, INT *los)
{
  (void)rnk_n; (void)iblock; (void)oblock;
  local_size_generic(ni, no, comm, pid, flags, 0, 0, lni, lis, lno, los);
}

void PX(local_block_many_dft_c2r)(int rnk_n, const INT *ni, const INT *no, const INT *iblock,
    const INT *oblock, MPI_Comm comm, int pid, unsigned flags,
    INT *lni, INT *lis, INT *lno, INT *los)
{
  (void)rnk_n; (void)iblock; (void)oblock;
  local_size_generic(ni, no, comm, pid, flags, 1, 0, lni, lis, lno, los);
}

INT PX(local_size_many_gc)(int rnk_n, const INT *local_n, const INT *local_n_start, INT howmany,
    const INT *gc_below, const INT *gc_above, INT *local_ngc, INT *local_gc_start)
{
  INT tot = howmany;
  for (int t = 0; t < rnk_n; t++) {
    local_ngc[t] = local_n[t] + gc_below[t] + gc_above[t];
    local_gc_start[t] = local_n_start[t] - gc_below[t];
    tot *= local_ngc[t];
  }
  return tot;
}

/* ------------------------------------------------------------------ */
/* host FFT (mixed radix, double precision)                            */
/* ------------------------------------------------------------------ */
typedef struct { INT n; Z *tw; } fft1d;

static fft1d *fft1d_new(INT n)
{
  fft1d *f = (fft1d *)malloc(sizeof(fft1d));
  f->n = n;
  f->tw = (Z *)malloc(sizeof(Z) * (size_t)n);
  for (INT j = 0; j < n; j++) {
    long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)j / (long double)n;
    f->tw[j] = (double)cosl(a) + (double)sinl(a) * I;
  }
  return f;
}
static void fft1d_free(fft1d *f) { if (f) { free(f->tw); free(f); } }

static INT smallest_factor(INT n)
{
  for (INT p = 2; p * p <= n; p++) if (n % p == 0) return p;
  return n;
}

/* out[0..n) = DFT of in[0], in[is], ... ; twiddle W_n^j = tw[j*ts] (conj if sign>0) */
static void fft_rec(const fft1d *f, INT n, INT ts, const Z *in, INT is, Z *out, int sign)
{
  if (n == 1) { out[0] = in[0]; return; }
  INT p = smallest_factor(n), m = n / p;
  for (INT r = 0; r < p; r++) fft_rec(f, m, ts * p, in + r * is, is * p, out + r * m, sign);
  Z tmp[64];
  Z *t = (p <= 64) ? tmp : (Z *)malloc(sizeof(Z) * (size_t)p);
  for (INT k = 0; k < m; k++) {
    for (INT r = 0; r < p; r++) {
      Z w = f->tw[((r * k) % n) * ts];
      if (sign > 0) w = conj(w);
      t[r] = out[r * m + k] * w;
    }
    if (p == 2) {
      out[k] = t[0] + t[1];
      out[k + m] = t[0] - t[1];
    } else {
      for (INT q = 0; q < p; q++) {
        Z s = 0;
        for (INT r = 0; r < p; r++) {
          Z w = f->tw[((r * q * m) % n) * ts];
          if (sign > 0) w = conj(w);
          s += t[r] * w;
        }
        out[q * m + k] = s;
      }
    }
  }
  if (t != tmp) free(t);
}

/* Pruned, shifted 1-d DFT of length n: in[i] <-> k = ik0 + i, out[o] <-> l = ok0 + o,
 * out[o] = sum_i in[i] exp(sign*2*pi*I*k*l/n).  Strides in units of Z. */
static void dft_pruned(const fft1d *f, int sign, INT ni, INT ik0, const Z *in, INT is,
                       INT no, INT ok0, Z *out, INT os, Z *buf, Z *buf2)
{
  INT n = f->n;
  for (INT j = 0; j < n; j++) buf[j] = 0;
  for (INT i = 0; i < ni; i++) {
    INT k = ik0 + i;
    buf[((k % n) + n) % n] += in[i * is];
  }
  fft_rec(f, n, 1, buf, 1, buf2, sign);
  for (INT o = 0; o < no; o++) {
    INT l = ok0 + o;
    out[o * os] = buf2[((l % n) + n) % n];
  }
}

/* ------------------------------------------------------------------ */
/* plans                                                               */
/* ------------------------------------------------------------------ */
enum { KIND_C2C = 0, KIND_C2R = 1, KIND_R2C = 2 };

struct PX(plan_s) {
  int kind, sign;
  unsigned flags;
  INT n[3], ni[3], no[3];
  INT lni[3], lis[3], lno[3], los[3];
  void *in, *out;
  MPI_Comm comm;
  fft1d *f[3];
};

static PX(plan) mkplan_generic(int kind, const INT *n, const INT *ni, const INT *no, void *in, void *out,
                               MPI_Comm comm, int sign, unsigned flags)
{
  PX(plan) p = (PX(plan))calloc(1, sizeof(*p));
  p->kind = kind; p->sign = sign; p->flags = flags; p->in = in; p->out = out; p->comm = comm;
  for (int t = 0; t < 3; t++) { p->n[t] = n[t]; p->ni[t] = ni[t]; p->no[t] = no[t]; p->f[t] = fft1d_new(n[t]); }
  if (!(flags & PFFT_SHIFTED_IN) || !(flags & PFFT_SHIFTED_OUT)) {
    fprintf(stderr, "shim PFFT: only SHIFTED_IN|SHIFTED_OUT plans are emulated\n"); abort();
  }
  local_size_generic(ni, no, comm, -1, flags, kind == KIND_C2R, kind == KIND_R2C, p->lni, p->lis, p->lno, p->los);
  return p;
}

PX(plan) PX(plan_many_dft)(int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
    const INT *iblock, const INT *oblock, C *in, C *out, MPI_Comm comm, int sign, unsigned flags)
{
  (void)rnk_n; (void)howmany; (void)iblock; (void)oblock;
  return mkplan_generic(KIND_C2C, n, ni, no, in, out, comm, sign, flags);
}
PX(plan) PX(plan_many_dft_c2r)(int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
    const INT *iblock, const INT *oblock, C *in, R *out, MPI_Comm comm, int sign, unsigned flags)
{
  (void)rnk_n; (void)howmany; (void)iblock; (void)oblock;
  return mkplan_generic(KIND_C2R, n, ni, no, in, out, comm, sign, flags);
}
PX(plan) PX(plan_many_dft_r2c)(int rnk_n, const INT *n, const INT *ni, const INT *no, INT howmany,
    const INT *iblock, const INT *oblock, R *in, C *out, MPI_Comm comm, int sign, unsigned flags)
{
  (void)rnk_n; (void)howmany; (void)iblock; (void)oblock;
  return mkplan_generic(KIND_R2C, n, ni, no, in, out, comm, sign, flags);
}

void PX(destroy_plan)(PX(plan) p)
{
  if (!p) return;
  for (int t = 0; t < 3; t++) fft1d_free(p->f[t]);
  free(p);
}

/* memory offset of logical element (i0,i1,i2) (local indices) in a local block */
static inline INT loc_off(const INT *ln, int transposed, INT i0, INT i1, INT i2)
{
  return transposed ? (i1 * ln[2] + i2) * ln[0] + i0 : (i0 * ln[1] + i1) * ln[2] + i2;
}

typedef struct {
  Z *gin, *t1, *t2, *gout;   /* shared global work arrays */
} exec_shared;

void PX(execute)(const PX(plan) p)
{
  MPI_Comm comm = p->comm;
  int rank, size;
  MPI_Comm_rank(comm, &rank); MPI_Comm_size(comm, &size);

  /* Extents / first index of the global logical arrays.  c2r: the input stores the half
   * spectrum k2 in [-ni2/2, 0]; the real output is defined as
   *     Re( sum_{k2=0} g e^{..} + 2 sum_{k2<0} g e^{..} )
   * i.e. every stored k2<0 entry also stands for its conjugate partner at -k (the convention
   * of the reference's direct transform, kernel/ndft-parallel.c:586-591, for Hermitian-
   * consistent input).  r2c: plain complex transform of the real input, output truncated to
   * k2 in [-no2/2, 0] (kernel/ndft-parallel.c:617-722). */
  INT ie[3], ik0[3], oe[3], ok0[3];
  for (int t = 0; t < 3; t++) { ie[t] = p->ni[t]; ik0[t] = -(p->ni[t] / 2); oe[t] = p->no[t]; ok0[t] = -(p->no[t] / 2); }
  if (p->kind == KIND_C2R) ie[2] = p->ni[2] / 2 + 1;
  if (p->kind == KIND_R2C) oe[2] = p->no[2] / 2 + 1;
  const int tin = (p->flags & PFFT_TRANSPOSED_IN) != 0, tout = (p->flags & PFFT_TRANSPOSED_OUT) != 0;

  exec_shared sh_local, *sh = &sh_local;
  if (rank == 0) {
    sh->gin = (Z *)calloc((size_t)(ie[0] * ie[1] * ie[2]), sizeof(Z));
    sh->t1 = (Z *)malloc(sizeof(Z) * (size_t)(ie[0] * ie[1] * oe[2]));
    sh->t2 = (Z *)malloc(sizeof(Z) * (size_t)(ie[0] * oe[1] * oe[2]));
    sh->gout = (Z *)malloc(sizeof(Z) * (size_t)(oe[0] * oe[1] * oe[2]));
  }
  void **tab = shim_publish(comm, sh);
  exec_shared S = *(exec_shared *)tab[0];
  shim_unpublish(comm);

  /* 1. copy local input block into the global input array */
  {
    const INT *ln = p->lni, *ls = p->lis;
    for (INT i0 = 0; i0 < ln[0]; i0++)
      for (INT i1 = 0; i1 < ln[1]; i1++)
        for (INT i2 = 0; i2 < ln[2]; i2++) {
          INT off = loc_off(ln, tin, i0, i1, i2);
          INT g0 = ls[0] + i0 - ik0[0], g1 = ls[1] + i1 - ik0[1], g2 = ls[2] + i2 - ik0[2];
          Z v = (p->kind == KIND_R2C) ? (Z)((const R *)p->in)[off] : (Z)((const C *)p->in)[off];
          if (p->kind == KIND_C2R && ls[2] + i2 < 0) v *= 2.0;
          S.gin[(g0 * ie[1] + g1) * ie[2] + g2] = v;
        }
  }
  MPI_Barrier(comm);

  Z *buf = (Z *)malloc(sizeof(Z) * (size_t)(2 * (p->n[0] + p->n[1] + p->n[2])));
  /* 2. dim 2 */
  for (INT r = rank; r < ie[0] * ie[1]; r += size)
    dft_pruned(p->f[2], p->sign, ie[2], ik0[2], S.gin + r * ie[2], 1, oe[2], ok0[2], S.t1 + r * oe[2], 1,
               buf, buf + p->n[2]);
  MPI_Barrier(comm);
  /* 3. dim 1 */
  for (INT r = rank; r < ie[0] * oe[2]; r += size) {
    INT i0 = r / oe[2], l2 = r % oe[2];
    dft_pruned(p->f[1], p->sign, ie[1], ik0[1], S.t1 + i0 * ie[1] * oe[2] + l2, oe[2],
               oe[1], ok0[1], S.t2 + i0 * oe[1] * oe[2] + l2, oe[2], buf, buf + p->n[1]);
  }
  MPI_Barrier(comm);
  /* 4. dim 0 */
  for (INT r = rank; r < oe[1] * oe[2]; r += size)
    dft_pruned(p->f[0], p->sign, ie[0], ik0[0], S.t2 + r, oe[1] * oe[2], oe[0], ok0[0], S.gout + r, oe[1] * oe[2],
               buf, buf + p->n[0]);
  free(buf);
  MPI_Barrier(comm);

  /* 5. copy my output block */
  {
    const INT *ln = p->lno, *ls = p->los;
    for (INT i0 = 0; i0 < ln[0]; i0++)
      for (INT i1 = 0; i1 < ln[1]; i1++)
        for (INT i2 = 0; i2 < ln[2]; i2++) {
          INT off = loc_off(ln, tout, i0, i1, i2);
          INT g0 = ls[0] + i0 - ok0[0], g1 = ls[1] + i1 - ok0[1], g2 = ls[2] + i2 - ok0[2];
          Z v = S.gout[(g0 * oe[1] + g1) * oe[2] + g2];
          if (p->kind == KIND_C2R) ((R *)p->out)[off] = (R)creal(v);
          else ((C *)p->out)[off] = (C)v;
        }
  }
  MPI_Barrier(comm);
  if (rank == 0) { free(S.gin); free(S.t1); free(S.t2); free(S.gout); }
}

/* ------------------------------------------------------------------ */
/* ghost cells                                                         */
/* ------------------------------------------------------------------ */
struct PX(gcplan_s) {
  int is_complex;
  INT n[3], ln[3], ls[3], below[3], above[3], lngc[3];
  void *data;
  MPI_Comm comm;
};

static PX(gcplan) mkgc(int is_complex, const INT *n, const INT *below, const INT *above, void *data, MPI_Comm comm)
{
  PX(gcplan) g = (PX(gcplan))calloc(1, sizeof(*g));
  int np[3], co[3];
  g->is_complex = is_complex; g->data = data; g->comm = comm;
  mesh_of(comm, -1, np, co);
  decompose(n, np, co, 0, 1, 0, g->ln, g->ls);
  for (int t = 0; t < 3; t++) {
    g->n[t] = n[t]; g->below[t] = below[t]; g->above[t] = above[t];
    g->lngc[t] = g->ln[t] + below[t] + above[t];
  }
  return g;
}

PX(gcplan) PX(plan_many_cgc)(int rnk_n, const INT *n, INT howmany, const INT *block, const INT *below,
    const INT *above, C *data, MPI_Comm comm, unsigned gc_flags)
{
  (void)rnk_n; (void)howmany; (void)block; (void)gc_flags;
  return mkgc(1, n, below, above, data, comm);
}
PX(gcplan) PX(plan_many_rgc)(int rnk_n, const INT *n, INT howmany, const INT *block, const INT *below,
    const INT *above, R *data, MPI_Comm comm, unsigned gc_flags)
{
  (void)rnk_n; (void)howmany; (void)block; (void)gc_flags;
  return mkgc(0, n, below, above, data, comm);
}
void PX(destroy_gcplan)(PX(gcplan) g) { free(g); }

static inline INT wrap(INT i, INT n) { i %= n; return i < 0 ? i + n : i; }

/* compact local block -> padded block with periodic halos */
void PX(exchange)(PX(gcplan) g)
{
  MPI_Comm comm = g->comm;
  int rank; MPI_Comm_rank(comm, &rank);
  const INT n0 = g->n[0], n1 = g->n[1], n2 = g->n[2];
  const int tup = g->is_complex ? 2 : 1;
  R *glob = NULL;
  if (rank == 0) glob = (R *)malloc(sizeof(R) * (size_t)(n0 * n1 * n2 * tup));
  void **tab = shim_publish(comm, glob);
  glob = (R *)tab[0];
  shim_unpublish(comm);

  const R *d = (const R *)g->data;
  for (INT i0 = 0; i0 < g->ln[0]; i0++)
    for (INT i1 = 0; i1 < g->ln[1]; i1++) {
      INT g0 = g->ls[0] + i0 + n0 / 2, g1 = g->ls[1] + i1 + n1 / 2;
      memcpy(glob + ((g0 * n1 + g1) * n2) * tup, d + ((i0 * g->ln[1] + i1) * g->ln[2]) * tup,
             sizeof(R) * (size_t)(g->ln[2] * tup));
    }
  MPI_Barrier(comm);
  R *o = (R *)g->data;
  for (INT i0 = 0; i0 < g->lngc[0]; i0++)
    for (INT i1 = 0; i1 < g->lngc[1]; i1++)
      for (INT i2 = 0; i2 < g->lngc[2]; i2++) {
        INT g0 = wrap(g->ls[0] - g->below[0] + i0 + n0 / 2, n0);
        INT g1 = wrap(g->ls[1] - g->below[1] + i1 + n1 / 2, n1);
        INT g2 = wrap(g->ls[2] - g->below[2] + i2 + n2 / 2, n2);
        INT src = ((g0 * n1 + g1) * n2 + g2) * tup, dst = ((i0 * g->lngc[1] + i1) * g->lngc[2] + i2) * tup;
        for (int c = 0; c < tup; c++) o[dst + c] = glob[src + c];
      }
  MPI_Barrier(comm);
  if (rank == 0) free(glob);
}

/* padded block -> halo contributions added to their owners, compact layout */
void PX(reduce)(PX(gcplan) g)
{
  MPI_Comm comm = g->comm;
  int rank, size; MPI_Comm_rank(comm, &rank); MPI_Comm_size(comm, &size);
  const INT n0 = g->n[0], n1 = g->n[1], n2 = g->n[2];
  const int tup = g->is_complex ? 2 : 1;
  R *glob = NULL;
  if (rank == 0) glob = (R *)calloc((size_t)(n0 * n1 * n2 * tup), sizeof(R));
  void **tab = shim_publish(comm, glob);
  glob = (R *)tab[0];
  shim_unpublish(comm);

  /* ranks add one after the other: deterministic summation order */
  for (int r = 0; r < size; r++) {
    if (r == rank) {
      const R *d = (const R *)g->data;
      for (INT i0 = 0; i0 < g->lngc[0]; i0++)
        for (INT i1 = 0; i1 < g->lngc[1]; i1++)
          for (INT i2 = 0; i2 < g->lngc[2]; i2++) {
            INT g0 = wrap(g->ls[0] - g->below[0] + i0 + n0 / 2, n0);
            INT g1 = wrap(g->ls[1] - g->below[1] + i1 + n1 / 2, n1);
            INT g2 = wrap(g->ls[2] - g->below[2] + i2 + n2 / 2, n2);
            INT dst = ((g0 * n1 + g1) * n2 + g2) * tup, src = ((i0 * g->lngc[1] + i1) * g->lngc[2] + i2) * tup;
            for (int c = 0; c < tup; c++) glob[dst + c] += d[src + c];
          }
    }
    MPI_Barrier(comm);
  }
  R *o = (R *)g->data;
  for (INT i0 = 0; i0 < g->ln[0]; i0++)
    for (INT i1 = 0; i1 < g->ln[1]; i1++) {
      INT g0 = g->ls[0] + i0 + n0 / 2, g1 = g->ls[1] + i1 + n1 / 2;
      memcpy(o + ((i0 * g->ln[1] + i1) * g->ln[2]) * tup, glob + ((g0 * n1 + g1) * n2) * tup,
             sizeof(R) * (size_t)(g->ln[2] * tup));
    }
  MPI_Barrier(comm);
  if (rank == 0) free(glob);
}

/* ------------------------------------------------------------------ */
/* test-data / printing helpers used by the reference API              */
/* ------------------------------------------------------------------ */
/* PFFT's own generator is not available; the formula is the one the reference's
 * tests/check_vs_pfft.c:167-181 uses for the same purpose ("parity unpinned"). */
void PX(init_input_complex_3d)(const INT *n, const INT *local_n, const INT *local_n_start, C *data)
{
  INT m = 0;
  for (INT k0 = local_n_start[0]; k0 < local_n_start[0] + local_n[0]; k0++)
    for (INT k1 = local_n_start[1]; k1 < local_n_start[1] + local_n[1]; k1++)
      for (INT k2 = local_n_start[2]; k2 < local_n_start[2] + local_n[2]; k2++, m++) {
        INT g = ((k0 + n[0] / 2) * n[1] + (k1 + n[1] / 2)) * n[2] + (k2 + n[2] / 2);
        data[m] = (R)(1000.0 / (2 * g + 1)) + (R)(1000.0 / (2 * g + 2)) * I;
      }
}

void PX(apr_complex_3d)(const C *data, const INT *local_n, const INT *local_n_start, const char *name, MPI_Comm comm)
{
  int rank; MPI_Comm_rank(comm, &rank);
  printf("Rank %d, %s: block %td x %td x %td at (%td,%td,%td), first = %.6e+%.6ei\n", rank, name,
         local_n[0], local_n[1], local_n[2], local_n_start[0], local_n_start[1], local_n_start[2],
         (double)creal(data[0]), (double)cimag(data[0]));
}
void PX(apr_real_3d)(const R *data, const INT *local_n, const INT *local_n_start, const char *name, MPI_Comm comm)
{
  int rank; MPI_Comm_rank(comm, &rank);
  printf("Rank %d, %s: block %td x %td x %td at (%td,%td,%td), first = %.6e\n", rank, name,
         local_n[0], local_n[1], local_n[2], local_n_start[0], local_n_start[1], local_n_start[2],
         (double)data[0]);
}

void PX(print_average_timer_adv)(const PX(plan) p, MPI_Comm comm) { (void)p; (void)comm; }
void PX(write_average_timer_adv)(const PX(plan) p, const char *name, MPI_Comm comm) { (void)p; (void)name; (void)comm; }
void PX(print_average_gctimer_adv)(const PX(gcplan) g, MPI_Comm comm) { (void)g; (void)comm; }
void PX(write_average_gctimer_adv)(const PX(gcplan) g, const char *name, MPI_Comm comm) { (void)g; (void)name; (void)comm; }
