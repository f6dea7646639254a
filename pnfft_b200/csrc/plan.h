// Host-side state behind the opaque pnfft_plan / pnfft_nodes handles
// (the reference's plan_s / nodes_s, kernel/ipnfft.h:160-256, re-thought for a device-resident pipeline).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <mpi.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "window.h"

namespace pnb {

typedef ptrdiff_t INT;

#define PNB_CUDA(call)                                                                            \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      fprintf(stderr, "pnfft-b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e__), __FILE__, \
              __LINE__, cudaGetErrorString(e__));                                                 \
      abort();                                                                                    \
    }                                                                                             \
  } while (0)

#define PNB_CUFFT(call)                                                                           \
  do {                                                                                            \
    cufftResult r__ = (call);                                                                     \
    if (r__ != CUFFT_SUCCESS) {                                                                   \
      fprintf(stderr, "pnfft-b200: cuFFT error %d at %s:%d\n", (int)r__, __FILE__, __LINE__);     \
      abort();                                                                                    \
    }                                                                                             \
  } while (0)

template <class R> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

// ---------------------------------------------------------------------------------------------
// Block decomposition (reference kernel/ndft-parallel.c:788-822 via PFFT's default blocks:
// block = ceil(n/P), rank c owns [c*block, min(n,(c+1)*block)), shifted by -n/2).
// ---------------------------------------------------------------------------------------------
inline void block_1d(INT n, int p, int c, INT *len, INT *start) {
  const INT blk = (n + p - 1) / p;
  INT s = (INT)c * blk;
  INT l = n - s;
  if (l > blk) l = blk;
  if (l <= 0) { l = 0; s = 0; }
  *len = l;
  *start = s;
}

struct Mesh {
  int np[2] = {1, 1};   // process mesh p0 x p1 over grid dims 0 and 1
  int co[2] = {0, 0};   // my coordinates
  int rank = 0, size = 1;
  int rank_of(int c0, int c1) const { return ((c0 % np[0] + np[0]) % np[0]) * np[1] + ((c1 % np[1] + np[1]) % np[1]); }
};

// Everything integer about a plan: sizes, local blocks, padded-grid geometry.
struct Layout {
  INT N[3], n[3], no[3];
  int m, cutoff;
  bool c2r;
  bool transposed;         // PNFFT_TRANSPOSED_F_HAT: f_hat block is local_N[1] x local_N[2] x local_N[0] in memory, k1 split over
                           // mesh dim 0, k2 over mesh dim 1, k0 whole (reference kernel/matrix_D.c:331-341, doc/intro.tex:39-44)
  INT Nc2;                 // stored extent of f_hat's last dim: N[2] or N[2]/2+1
  INT local_N[3], local_N_start[3];
  INT local_no[3], local_no_start[3];
  INT gcb[3], gca[3];      // ghost cells below / above (reference get_size_gcells :1550-1561)
  INT ngc[3];              // padded extents local_no + gcb + gca (reference local_array_size :1575-1582)
  INT pitch2;              // row pitch (cells) of the device padded grid, >= ngc[2], 16-byte aligned rows
  INT o_off[3];            // first kept index of the length-n FFT output (shifted storage): n/2 - no/2
};

// ---------------------------------------------------------------------------------------------
// Device-side description handed to the gridding kernels by value
// ---------------------------------------------------------------------------------------------
template <class R> struct GridGeom {
  int m, cutoff;
  int kind;                // WindowKind
  int fast_gauss;
  R n[3];                  // (R) n_t
  R b[3];
  int los[3];              // local_no_start
  int lno[3];              // local_no
  int ngc[3];              // padded extents
  long long pitch1;        // = pitch2 (cells per y-row)
  long long pitch0;        // = ngc[1] * pitch2 (cells per x-plane)
  const R *exp_const;      // fast Gaussian table [3][cutoff] (device)
  const R *poly;           // per-tap window polynomials [deg+1][3*cutoff] in u = 2 frac - 1 (device) or nullptr
  int poly_deg;
  int poly_deg_psi;        // degree that suffices for psi alone (<= poly_deg)
  // PNFFT_INTERLACED (reference kernel/ndft-parallel.c:2732-2753): in the second pass every node is shifted by half a mesh
  // width, x_t += 0.5 / n_t, and folded back into [-0.5, 0.5) AFTER its grid index was taken; both passes weigh by 0.5
  // (kernel/assign.c:481,689), which the kernels fold into the x-axis window factors (a power of two: exact)
  // PNFFT_PRE_{CONST,LIN,QUAD,CUB}_PSI (reference kernel/ndft-parallel.c:321-353, 1586-1617): window values looked up in
  // tables of intpol_num nodes per grid interval and interpolated with order intpol_order (-1: direct evaluation);
  // table of axis t, derivative q: intpol_tab[3 * q + t], laid out [node k][tap c][stencil point i]
  int intpol_order;
  int intpol_num;
  const R *intpol_tab[9];
  int il_on;               // this pass shifts the nodes
  double il[3];            // 0.5 / n_t
  R wscale;                // 1, or 0.5 for the two passes of an interlaced plan
};


// ---------------------------------------------------------------------------------------------
// Strided 3-d box map (see boxcopy.cuh) and the description of the pencil-FFT re-distributions
// (see fftpipe.cuh)
// ---------------------------------------------------------------------------------------------
struct BoxMap {
  long long dims[3];
  long long a_off, a_str[3];
  long long c_off, c_str[3];
  int parity;   // used by the sign variant: factor (-1)^(i0+i1+i2+parity)
};

struct Transfer {
  int peer;
  long long send_elems = 0, recv_elems = 0;  // complex elements
  long long send_off = 0, recv_off = 0;      // offsets inside the chunk buffers
  std::vector<BoxMap> send_maps, recv_maps;  // source array <-> send chunk ; destination array <-> recv chunk
  std::vector<BoxMap> self_maps;             // peer == me: destination array (a side) <-> source array (c side), no chunk
  bool send_sign = false;
};

struct Stage {
  std::vector<Transfer> tr;
  long long src_elems = 0, dst_elems = 0, send_total = 0, recv_total = 0;
  long long zero_off = 0, zero_len = 0;      // contiguous gap of the destination array (forward direction)
  bool zero_all = false;
};

struct PipeGeom {
  Stage st[3];
  long long L1_elems, L3_elems, L4_elems;   // complex elements
  long long S1, S3;                          // strides of the x / y passes
  INT z0len, z1len, n2z;                     // n2z: complex row length of L4 (n2 or n2/2+1)
  long long buf_elems;                       // complex elements each work buffer needs
};

// Peer-memory data plane of a multi-rank plan (p2p.cuh): every other rank's work buffers, padded grid, g1 and flag array
// as mapped into this process with CUDA IPC, and the device-resident job lists of the exchange stages.
struct JobList;
struct PeerPlane {
  bool on = false;
  int nranks = 1, rank = 0;
  std::vector<void *> work[3];     // [rank] base of d_work[k] as mapped here
  std::vector<void *> grid;        // [rank] padded grid
  std::vector<void *> g1;          // [rank] compact FFT-input-side array
  std::vector<int *> flags;        // [rank] flag array (nranks ints)
  int **d_flag_ptrs = nullptr;     // device copy of flags[]
  int *d_error = nullptr, *h_error = nullptr;
  int epoch = 0;
  JobList *d_jobs = nullptr;       // device job lists (LIST_* in core.cuh)
  int nblocks[16] = {0};           // blocks of each list (0: nothing to do)
};

template <class R> struct Nodes;

template <class R> struct Plan {
  typedef typename Vec2<R>::type C;

  Layout L;
  Mesh mesh;
  MPI_Comm comm = MPI_COMM_NULL;
  unsigned pnfft_flags = 0, pfft_flags = 0;
  R x_max[3], sigma[3], b[3];
  int kind = WIN_KAISER_BESSEL;
  unsigned win_gen = 0;          // bumped whenever the window shape b changes (pnfft_set_b): cached node-table rows carry it

  // user-visible f_hat (host or device pointer); owned iff PNFFT_MALLOC_F_HAT
  C *f_hat = nullptr;
  bool owns_f_hat = false;

  // device state
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // second H2D queue: node coordinates travel while D and F run (Core::trafo)
  cudaEvent_t ev_copy[3] = {nullptr, nullptr, nullptr};
  cudaStream_t node_stream = nullptr;   // trafo: binning + node table run here while D, F and the halo exchange run on `stream`
  int b_phase = 3;                      // what launch_B does: bit 0 = node table, bit 1 = gridding kernel (z-march v2 only)
  bool side_nodes = false;              // the current trafo ran its node side on node_stream (stage timers: ev[9..12])
  bool x_first = false;                 // trafo: coordinates uploaded before f_hat (Core::trafo)
  bool x_via_copy_stream = false;       // set by trafo around prepare_nodes: ev_copy[0] marks where the x upload may start
  C *d_f_hat = nullptr;          // staging copy when the user's f_hat is a host pointer
  R *d_invphi[3] = {nullptr, nullptr, nullptr};  // 1/phi_hat tables incl. the (-1)^k fft-shift sign, [local_N[t]]
  R *d_exp_const = nullptr;      // [3][cutoff] for FAST_GAUSSIAN
  R *d_poly = nullptr;           // window polynomials (see GridGeom::poly)
  int intpol_order = -1;         // PNFFT_PRE_*_PSI interpolation order, -1: none
  int intpol_num = 0;
  R *d_intpol[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int poly_deg = -1, poly_deg_psi = -1;
  int use_poly = 1;
  void *d_grid = nullptr;        // padded grid [ngc0][ngc1][pitch2] of C (c2c) or R (c2r)
  void *d_work[3] = {nullptr, nullptr, nullptr};  // ping-pong FFT stage buffers + the pack buffer of the re-distributions
  size_t work_bytes = 0, grid_bytes = 0;
  C *d_g1 = nullptr;             // compact FFT-input-side array (local_N block) used with OMIT_* flags / ik
  C *d_g1_buffer = nullptr;      // ik differentiation buffer

  PipeGeom pipe;
  PeerPlane peer;
  // FFT plans
  cufftHandle fft_x = 0, fft_y = 0, fft_z_fwd = 0, fft_z_bwd = 0;
  bool fft_ready = false;
  // own FFT kernels (fftown.cuh) per axis: power-of-two complex lengths; the twiddle tables exp(-2 pi i q / n)
  bool own_fft[3] = {false, false, false};
  void *d_tw[3] = {nullptr, nullptr, nullptr};

  // timers (seconds), reference slots api/pnfft.h:407-418
  double timer_trafo[10], timer_adj[10];
  double stage_ms[2][8];
  long long launches = 0;       // kernels of this library launched so far
  long long lib_launches = 0;   // cuFFT / CUB / NCCL calls issued so far
  int kernel_variant = 0;
  int il_pass = 0;              // which pass of an interlaced transform is being issued (Core::geom)
  bool warned_hessian = false;

  cudaEvent_t ev[16];

  // scratch for binning (grown on demand)
  void *d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
};

// What binning (and PNFFT_PRE_PSI) derive from the node coordinates.  An interlaced plan keeps two of these: one for the
// nodes as given, one for the nodes shifted by half a mesh width (reference nodes->pre_psi / pre_psi_il).
template <class R> struct BinState {
  int *d_tile = nullptr;       // tile id per node
  int *d_tile_sorted = nullptr;
  int *d_perm = nullptr;       // sorted position -> node index
  int *d_idx = nullptr;        // identity
  int *d_tile_count = nullptr; // [ntiles+1]
  int *d_tile_start = nullptr; // [ntiles+1]
  size_t cap_nodes = 0, cap_tiles = 0;
  bool binned = false;         // valid for the current x (kept across calls after precompute_psi or while x is declared static)
  const void *bin_plan = nullptr;   // plan whose geometry the bins were made for
  int bin_family = -1;              // ... and its kernel family (tile shape)
  unsigned long long bin_hash = 0;  // content hash of the device-resident x the bins were made from (0: none)
  const R *d_x_bound = nullptr; // device x the binning / tables were computed from
  int *h_maxcol = nullptr;     // pinned: largest column population seen by the last finished binning (load-balance hint)
  int *d_maxcol = nullptr;
  // PNFFT_PRE_PSI tables in sorted order (reference kernel/ndft-parallel.c:1184-1240)
  R *d_pre_psi = nullptr, *d_pre_dpsi = nullptr;
  // node-table rows of the tensor-core gridding kernels (zmarch3.cuh), kept while this binning stays valid: the window
  // factors depend on x alone, so trafo and adj (and every further call on the same coordinates) share one table
  R *d_rows = nullptr;
  size_t cap_rows = 0;
  int rows_flavor = -1;            // -1: none, 0: psi sections only, 1: with the derivative sections
  const void *rows_plan = nullptr;
  unsigned rows_gen = 0;           // Plan::win_gen the rows were evaluated with
};

template <class R> struct Nodes : BinState<R> {
  typedef typename Vec2<R>::type C;
  INT local_M = 0;
  unsigned malloc_flags = 0;
  R *x = nullptr;
  R *f = nullptr;          // C[M] (c2c) or R[M] (c2r), user pointer (host or device)
  R *grad_f = nullptr;     // 3 per node
  R *hessian_f = nullptr;  // 6 per node (xx, xy, xz, yy, yz, zz), trafo only
  unsigned precompute_flags = 0;

  // device mirrors (allocated when the user arrays live on the host)
  R *d_x = nullptr, *d_f = nullptr, *d_grad_f = nullptr, *d_hess = nullptr;
  size_t cap_x = 0, cap_f = 0, cap_grad = 0, cap_hess = 0;

  BinState<R> il;          // the same for the shifted nodes of the second interlacing pass
  void swap_il() { BinState<R> t = *static_cast<BinState<R> *>(this); *static_cast<BinState<R> *>(this) = il; il = t; }

  // "x has not changed since the last call" (extension pnfft_b200_nodes_x_static / PNFFT_B200_X_STATIC):
  // the uploaded coordinates and their binning are reused until pnfft_set_x / pnfft_precompute_psi / the switch say otherwise
  bool x_static = false;
  bool x_uploaded = false;     // d_x holds the current host x
  unsigned long long x_hash = 0;   // device-resident x: content hash the binning belongs to (0 = none)
  unsigned long long *d_hash = nullptr, *h_hash = nullptr;
  // pnfft_adj on host-resident coordinates that turned out unchanged last time (trafo then adj of one step): the next adj
  // runs on the bins it has while the coordinates travel into d_x_alt and are hashed on the copy stream; a different hash
  // makes d_x_alt the current mirror and the adjoint is redone (Core::adj)
  R *d_x_alt = nullptr;
  size_t cap_x_alt = 0;
  bool adj_same_x_last = false;
  bool spec_pending = false;
  unsigned long long known_hash = 0;   // hash of the coordinates handed to prepare_nodes as dx_known (0: none)

  // per-call node table of the z-marching kernels (zmarch.cuh: ZmTab), grown on demand
  R *d_wtab = nullptr;
  size_t cap_wtab = 0;
  // node values in sorted order for the tensor-core scatter (k_pack_vals)
  R *d_vals = nullptr;
  size_t cap_vals = 0;
};

}  // namespace pnb
