// Pruned, shifted pencil FFT on the device (the F / F^H matrices, reference
// kernel/ndft-parallel.c:1524-1546 = one pfft_execute of the plans made at :963-986) and the
// ghost-cell exchange / reduce that PFFT provides (reference call sites :2558, :2679).
//
// Design (B200-first, not PFFT's schedule): the three 1-d passes run in the order x, y, z so that
// every re-distribution moves the *pruned* (not yet zero-padded) data:
//   g1 [N0/p0][N1/p1][Z]  --E1(p0 group)-->  L1 [n0][N1/p1][Z/p0]      FFT along x (axis outermost)
//   L1                     --E2(all ranks)-->  L3 [n1][no0/p0][Z/p1]     FFT along y (axis outermost)
//   L3                     --E3(p1 group)-->  L4 [no0/p0][no1/p1][n2]   FFT along z (contiguous)
//   L4 --crop/embed--> padded grid [no0/p0+2m][no1/p1+2m][no2+2m] + halos
// E1..E3 move 1x, 2x and 4x the f_hat volume (sigma=2) instead of the 2x/4x/8x/8x of a z-y-x schedule.
// Each E is pack (box copies) -> NCCL send/recv all-to-all over NVLink -> unpack (box copies); with one
// rank the exchange degenerates to a device copy.  The adjoint runs the same description backwards.
// Index shifts (k in [-N/2,N/2), l in [-n/2,n/2)) are realised as k -> k mod n placement plus the
// factor (-1)^(k0+k1+k2) on the spectrum side (reference doc/manual.tex:178-238 uses twiddles too).
#pragma once
#include <nccl.h>

#include <algorithm>

#include "boxcopy.cuh"

namespace pnb {

ncclComm_t world_nccl();  // comm.cu: lazily created NCCL communicator over all ranks

// NCCL is bound at run time (dlopen of libnccl.so.2, comm.cu) the first time a multi-rank plan needs it: the library
// then shares whatever NCCL the process already loaded (e.g. the one bundled with PyTorch) instead of forcing a
// second copy into the process, and single-rank users need no NCCL at all.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
};
const NcclApi &nccl_api();

#define PNB_NCCL(call)                                                                             \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) {                                                                      \
      fprintf(stderr, "pnfft-b200: NCCL error %s at %s:%d\n", nccl_api().GetErrorString(r__), __FILE__, __LINE__); \
      abort();                                                                                     \
    }                                                                                              \
  } while (0)

struct Split {
  std::vector<INT> start, len;
  Split() {}
  Split(INT n, int p) : start((size_t)p), len((size_t)p) {
    for (int c = 0; c < p; c++) block_1d(n, p, c, &len[(size_t)c], &start[(size_t)c]);
  }
};

inline int pmod2(long long v) { return (int)(((v % 2) + 2) % 2); }

// add the (up to two) segments of  k in [k_start, k_start+len) -> position k mod n
template <class F> inline void wrap_segments(INT k_start, INT len, INT n, F f) {
  // f(i_begin, count, pos_begin)
  INT neg = 0;
  if (k_start < 0) neg = std::min(len, -k_start);
  if (neg > 0) f((INT)0, neg, n + k_start);
  if (len - neg > 0) f(neg, len - neg, k_start + neg);
}

// A rank's own chunk needs no chunk: compose "source array <-> dense send chunk" with "destination array <-> sub-box of
// that chunk" into one map destination array (a side) <-> source array (c side).  False if the sub-box does not follow
// the dense chunk's axes (the caller then keeps pack + unpack).
inline bool compose_self_map(const BoxMap &sm, const BoxMap &rm, BoxMap *out) {
  if (sm.c_str[2] != 1 || sm.c_str[1] != sm.dims[2] || sm.c_str[0] != sm.dims[1] * sm.dims[2]) return false;
  long long rel = rm.c_off - sm.c_off;
  if (rel < 0) return false;
  long long o[3];
  o[0] = sm.c_str[0] > 0 ? rel / sm.c_str[0] : 0; rel -= o[0] * sm.c_str[0];
  o[1] = sm.c_str[1] > 0 ? rel / sm.c_str[1] : 0; rel -= o[1] * sm.c_str[1];
  o[2] = rel;
  *out = rm;
  bool used[3] = {false, false, false};
  for (int k = 0; k < 3; k++) {
    int pick = -1;
    for (int d = 0; d < 3 && pick < 0; d++)
      if (!used[d] && sm.c_str[d] == rm.c_str[k] && o[d] + rm.dims[k] <= sm.dims[d]) pick = d;
    if (pick < 0) {
      if (rm.dims[k] != 1) return false;
      out->c_str[k] = 0;
      continue;
    }
    used[pick] = true;
    out->c_str[k] = sm.a_str[pick];
  }
  for (int d = 0; d < 3; d++) if (o[d] >= sm.dims[d]) return false;
  out->c_off = sm.a_off + o[0] * sm.a_str[0] + o[1] * sm.a_str[1] + o[2] * sm.a_str[2];
  out->parity = (int)((sm.parity + o[0] + o[1] + o[2]) & 1);
  return true;
}

inline PipeGeom build_pipe(const Layout &L, const Mesh &M) {
  PipeGeom G;
  const int p0 = M.np[0], p1 = M.np[1], a = M.co[0], b = M.co[1];
  const INT Zc = L.Nc2;
  Split SX(L.N[0], p0), SY(L.N[1], p1), Z0(Zc, p0), Z1(Zc, p1), XO(L.no[0], p0), YO(L.no[1], p1);
  const INT lN0 = L.local_N[0], lN1 = L.local_N[1];
  const INT lno0 = L.local_no[0], lno1 = L.local_no[1];
  const INT z0len = Z0.len[(size_t)a], z1len = Z1.len[(size_t)b];
  G.z0len = z0len; G.z1len = z1len;
  G.n2z = L.c2r ? L.n[2] / 2 + 1 : L.n[2];
  G.S1 = (long long)lN1 * z0len;
  G.S3 = (long long)lno0 * z1len;
  G.L1_elems = (long long)L.n[0] * G.S1;
  G.L3_elems = (long long)L.n[1] * G.S3;
  G.L4_elems = (long long)lno0 * lno1 * G.n2z;

  // ---- stage 1: g1 -> L1, exchange inside the p0 group (ranks (a', b)) ----
  if (!L.transposed) {
    Stage &S = G.st[0];
    S.src_elems = (long long)lN0 * lN1 * Zc;
    S.dst_elems = G.L1_elems;
    for (int ap = 0; ap < p0; ap++) {
      Transfer T;
      T.peer = M.rank_of(ap, b);
      T.send_sign = true;
      // what I send to a': my (k0,k1) block, z in Z0[a']
      const INT zl = Z0.len[(size_t)ap], zs = Z0.start[(size_t)ap];
      T.send_elems = (long long)lN0 * lN1 * zl;
      if (T.send_elems > 0) {
        BoxMap bm = dense_map(lN0, lN1, zl, lN1, Zc, 0, 0, zs, 0);
        bm.parity = pmod2((long long)L.local_N_start[0] + L.local_N_start[1] + (zs - L.N[2] / 2));
        T.send_maps.push_back(bm);
      }
      // what I receive from a': its k0 block, my k1 block, z in Z0[a]
      const INT xl = SX.len[(size_t)ap], xs = SX.start[(size_t)ap] - L.N[0] / 2;
      T.recv_elems = (long long)xl * lN1 * z0len;
      if (T.recv_elems > 0)
        wrap_segments(xs, xl, L.n[0], [&](INT ib, INT cnt, INT pos) {
          BoxMap bm = dense_map(cnt, lN1, z0len, lN1, z0len, pos, 0, 0, (long long)ib * lN1 * z0len);
          T.recv_maps.push_back(bm);
        });
      S.tr.push_back(T);
    }
    const INT hi = L.N[0] - L.N[0] / 2, lo = L.n[0] - L.N[0] / 2;  // rows [hi, lo) stay zero
    S.zero_off = (long long)hi * G.S1;
    S.zero_len = (long long)(lo - hi) * G.S1;
  }
  // ---- stage 2: L1 -> L3, exchange between all ranks ----
  if (!L.transposed) {
    Stage &S = G.st[1];
    S.src_elems = G.L1_elems;
    S.dst_elems = G.L3_elems;
    for (int ap = 0; ap < p0; ap++)
      for (int bp = 0; bp < p1; bp++) {
        Transfer T;
        T.peer = M.rank_of(ap, bp);
        {  // send to (a',b'): rows x in XO[a'] (storage offset o_off), my y block, z in Z0[a] ^ Z1[b']
          const INT zs = std::max(Z0.start[(size_t)a], Z1.start[(size_t)bp]);
          const INT ze = std::min(Z0.start[(size_t)a] + z0len, Z1.start[(size_t)bp] + Z1.len[(size_t)bp]);
          const INT zl = ze > zs ? ze - zs : 0;
          const INT xl = XO.len[(size_t)ap];
          T.send_elems = (long long)xl * lN1 * zl;
          if (T.send_elems > 0)
            T.send_maps.push_back(dense_map(xl, lN1, zl, lN1, z0len, XO.start[(size_t)ap] + L.o_off[0], 0,
                                            zs - Z0.start[(size_t)a], 0));
        }
        {  // receive from (a',b'): my x rows, its y block, z in Z0[a'] ^ Z1[b]
          const INT zs = std::max(Z0.start[(size_t)ap], Z1.start[(size_t)b]);
          const INT ze = std::min(Z0.start[(size_t)ap] + Z0.len[(size_t)ap], Z1.start[(size_t)b] + z1len);
          const INT zl = ze > zs ? ze - zs : 0;
          const INT yl = SY.len[(size_t)bp], ys = SY.start[(size_t)bp] - L.N[1] / 2;
          T.recv_elems = (long long)lno0 * yl * zl;
          if (T.recv_elems > 0)
            wrap_segments(ys, yl, L.n[1], [&](INT ib, INT cnt, INT pos) {
              BoxMap bm;
              bm.dims[0] = lno0; bm.dims[1] = cnt; bm.dims[2] = zl;
              bm.a_str[0] = z1len; bm.a_str[1] = (long long)lno0 * z1len; bm.a_str[2] = 1;
              bm.a_off = (long long)pos * lno0 * z1len + (zs - Z1.start[(size_t)b]);
              bm.c_str[0] = (long long)yl * zl; bm.c_str[1] = zl; bm.c_str[2] = 1;
              bm.c_off = (long long)ib * zl;
              bm.parity = 0;
              T.recv_maps.push_back(bm);
            });
        }
        S.tr.push_back(T);
      }
    const INT hi = L.N[1] - L.N[1] / 2, lo = L.n[1] - L.N[1] / 2;
    S.zero_off = (long long)hi * G.S3;
    S.zero_len = (long long)(lo - hi) * G.S3;
  }
  // ---- PNFFT_TRANSPOSED_F_HAT (reference kernel/matrix_D.c:331-341, ndft-parallel.c:965-978): the block already holds
  // every k0 of its (k1, k2) range, memory order [k1][k2][k0], so the x pass needs NO exchange and the y pass only one
  // inside the p0 group:
  //   g1 [N1/p0][Z/p1][N0]  --local-->  L1 [n0][N1/p0][Z/p1]  --E2(p0 group)-->  L3 [n1][no0/p0][Z/p1]
  if (L.transposed) {
    const INT lN1t = L.local_N[1], lN2t = L.local_N[2];
    Split SYt(L.N[1], p0);
    G.S1 = (long long)lN1t * lN2t;
    G.L1_elems = (long long)L.n[0] * G.S1;
    {
      Stage &S = G.st[0];
      S.src_elems = (long long)L.N[0] * G.S1;
      S.dst_elems = G.L1_elems;
      Transfer T;
      T.peer = M.rank;
      T.send_sign = true;
      T.send_elems = T.recv_elems = S.src_elems;
      if (S.src_elems > 0)
        wrap_segments(-(L.N[0] / 2), L.N[0], L.n[0], [&](INT ib, INT cnt, INT pos) {
          BoxMap bm;      // a side: L1 (destination), c side: g1 (source)
          bm.dims[0] = cnt; bm.dims[1] = lN1t; bm.dims[2] = lN2t;
          bm.a_off = (long long)pos * G.S1; bm.a_str[0] = G.S1; bm.a_str[1] = lN2t; bm.a_str[2] = 1;
          bm.c_off = ib; bm.c_str[0] = 1; bm.c_str[1] = (long long)lN2t * L.N[0]; bm.c_str[2] = L.N[0];
          bm.parity = pmod2((long long)(ib - L.N[0] / 2) + L.local_N_start[1] + L.local_N_start[2]);
          T.self_maps.push_back(bm);
        });
      S.tr.push_back(T);
      const INT hi = L.N[0] - L.N[0] / 2, lo = L.n[0] - L.N[0] / 2;
      S.zero_off = (long long)hi * G.S1;
      S.zero_len = (long long)(lo - hi) * G.S1;
    }
    {
      Stage &S = G.st[1];
      S.src_elems = G.L1_elems;
      S.dst_elems = G.L3_elems;
      for (int ap = 0; ap < p0; ap++) {
        Transfer T;
        T.peer = M.rank_of(ap, b);
        {  // send to a': rows x in XO[a'] (storage offset o_off), my k1 block, my z block
          const INT xl = XO.len[(size_t)ap];
          T.send_elems = (long long)xl * lN1t * lN2t;
          if (T.send_elems > 0)
            T.send_maps.push_back(dense_map(xl, lN1t, lN2t, lN1t, lN2t, XO.start[(size_t)ap] + L.o_off[0], 0, 0, 0));
        }
        {  // receive from a': my x rows, its k1 block, my z block -> position k1 mod n1
          const INT yl = SYt.len[(size_t)ap], ys = SYt.start[(size_t)ap] - L.N[1] / 2;
          T.recv_elems = (long long)lno0 * yl * z1len;
          if (T.recv_elems > 0)
            wrap_segments(ys, yl, L.n[1], [&](INT ib, INT cnt, INT pos) {
              BoxMap bm;
              bm.dims[0] = lno0; bm.dims[1] = cnt; bm.dims[2] = z1len;
              bm.a_str[0] = z1len; bm.a_str[1] = (long long)lno0 * z1len; bm.a_str[2] = 1;
              bm.a_off = (long long)pos * lno0 * z1len;
              bm.c_str[0] = (long long)yl * z1len; bm.c_str[1] = z1len; bm.c_str[2] = 1;
              bm.c_off = (long long)ib * z1len;
              bm.parity = 0;
              T.recv_maps.push_back(bm);
            });
        }
        S.tr.push_back(T);
      }
      const INT hi = L.N[1] - L.N[1] / 2, lo = L.n[1] - L.N[1] / 2;
      S.zero_off = (long long)hi * G.S3;
      S.zero_len = (long long)(lo - hi) * G.S3;
    }
  }
  // ---- stage 3: L3 -> L4, exchange inside the p1 group (ranks (a, b')) ----
  {
    Stage &S = G.st[2];
    S.src_elems = G.L3_elems;
    S.dst_elems = G.L4_elems;
    S.zero_all = true;
    for (int bp = 0; bp < p1; bp++) {
      Transfer T;
      T.peer = M.rank_of(a, bp);
      {  // send to b': rows y in YO[b'] (+o_off), all my x, my z block
        const INT yl = YO.len[(size_t)bp];
        T.send_elems = (long long)yl * lno0 * z1len;
        if (T.send_elems > 0)
          T.send_maps.push_back(dense_map(yl, lno0, z1len, lno0, z1len, YO.start[(size_t)bp] + L.o_off[1], 0, 0, 0));
      }
      {  // receive from b': chunk [lno1][lno0][Z1[b']] -> L4[x][y][zpos]
        const INT zl = Z1.len[(size_t)bp], zs = Z1.start[(size_t)bp];
        T.recv_elems = (long long)lno1 * lno0 * zl;
        if (T.recv_elems > 0) {
          if (!L.c2r) {
            wrap_segments(zs - L.N[2] / 2, zl, L.n[2], [&](INT ib, INT cnt, INT pos) {
              BoxMap bm;
              bm.dims[0] = lno1; bm.dims[1] = lno0; bm.dims[2] = cnt;
              bm.a_str[0] = G.n2z; bm.a_str[1] = (long long)lno1 * G.n2z; bm.a_str[2] = 1;
              bm.a_off = pos;
              bm.c_str[0] = (long long)lno0 * zl; bm.c_str[1] = zl; bm.c_str[2] = 1;
              bm.c_off = ib;
              bm.parity = 0;
              T.recv_maps.push_back(bm);
            });
          } else {
            // stored index i2 <-> k2 = i2 - N2/2 in [-N2/2, 0]; position j = -k2 (Hermitian half of the c2r pass)
            BoxMap bm;
            bm.dims[0] = lno1; bm.dims[1] = lno0; bm.dims[2] = zl;
            bm.a_str[0] = G.n2z; bm.a_str[1] = (long long)lno1 * G.n2z; bm.a_str[2] = -1;
            bm.a_off = L.N[2] / 2 - zs;
            bm.c_str[0] = (long long)lno0 * zl; bm.c_str[1] = zl; bm.c_str[2] = 1;
            bm.c_off = 0;
            bm.parity = 0;
            T.recv_maps.push_back(bm);
          }
        }
      }
      S.tr.push_back(T);
    }
  }
  long long need = std::max(std::max(G.L1_elems, G.L3_elems), G.L4_elems);
  static const bool no_self = getenv("PNFFT_B200_NO_SELF_MAPS") && atoi(getenv("PNFFT_B200_NO_SELF_MAPS")) != 0;
  for (int s = 0; s < 3; s++) {
    Stage &S = G.st[s];
    for (auto &T : S.tr) {
      if (!T.self_maps.empty()) continue;     // built directly (transposed f_hat, stage 1)
      if (no_self || T.peer != M.rank || T.send_maps.size() != 1 || T.recv_maps.empty() || T.send_elems != T.recv_elems) continue;
      std::vector<BoxMap> sm;
      bool ok = true;
      for (const auto &rm : T.recv_maps) {
        BoxMap c;
        ok = ok && compose_self_map(T.send_maps[0], rm, &c);
        if (ok) sm.push_back(c);
      }
      if (ok) T.self_maps = sm;
    }
    long long so = 0, ro = 0;
    for (auto &T : S.tr) { T.send_off = so; T.recv_off = ro; so += T.send_elems; ro += T.recv_elems; }
    S.send_total = so; S.recv_total = ro;
    need = std::max(need, std::max(so, ro));
  }
  // the real z output of c2r needs n2 reals per row; complex rows of n2/2+1 already cover that
  G.buf_elems = need + 16;
  return G;
}

// -----------------------------------------------------------------------------------------------
// exchange of the chunk buffers
// -----------------------------------------------------------------------------------------------
template <class C>
inline void exchange_chunks(const Stage &S, const Mesh &M, C *from, C *to, bool forward, cudaStream_t st, bool copy_self = true) {
  const bool multi = M.size > 1;
  if (multi) PNB_NCCL(nccl_api().GroupStart());
  for (const auto &T : S.tr) {
    const long long ns = forward ? T.send_elems : T.recv_elems, nr = forward ? T.recv_elems : T.send_elems;
    const long long os = forward ? T.send_off : T.recv_off, orr = forward ? T.recv_off : T.send_off;
    if (T.peer == M.rank) {
      if (copy_self && ns > 0) PNB_CUDA(cudaMemcpyAsync(to + orr, from + os, sizeof(C) * (size_t)ns, cudaMemcpyDeviceToDevice, st));
    } else {
      if (ns > 0) PNB_NCCL(nccl_api().Send(from + os, (size_t)ns * sizeof(C), ncclChar, T.peer, world_nccl(), st));
      if (nr > 0) PNB_NCCL(nccl_api().Recv(to + orr, (size_t)nr * sizeof(C), ncclChar, T.peer, world_nccl(), st));
    }
  }
  if (multi) PNB_NCCL(nccl_api().GroupEnd());
}

// forward: src array -> (pack) pk -> (exchange) w0 -> (unpack) w1 = the destination array.  src may alias w0 (it is dead
// once the exchange starts).  The chunk a rank sends to itself is one strided copy src -> w1 (Transfer::self_maps), or,
// where the maps do not compose, is unpacked straight from the pack buffer: with one rank a re-distribution is a single
// pass over the data.
template <class C>
inline void run_stage_forward(const Stage &S, const Mesh &M, C *src, C *w0, C *w1, C *pk, cudaStream_t st, long long *launches) {
  for (const auto &T : S.tr) {
    if (!T.self_maps.empty()) continue;
    for (const auto &bm0 : T.send_maps) {
      BoxMap bm = bm0; bm.c_off += T.send_off;
      box_copy<C>(st, src, pk, bm, BOX_A2C, T.send_sign, launches);
    }
  }
  if (S.zero_all) PNB_CUDA(cudaMemsetAsync(w1, 0, sizeof(C) * (size_t)S.dst_elems, st));
  else if (S.zero_len > 0) PNB_CUDA(cudaMemsetAsync(w1 + S.zero_off, 0, sizeof(C) * (size_t)S.zero_len, st));
  for (const auto &T : S.tr)
    for (const auto &bm : T.self_maps) box_copy<C>(st, w1, src, bm, BOX_C2A, T.send_sign, launches);
  exchange_chunks<C>(S, M, pk, w0, true, st, false);
  for (const auto &T : S.tr) {
    if (!T.self_maps.empty()) continue;
    const bool self = T.peer == M.rank;
    for (const auto &bm0 : T.recv_maps) {
      BoxMap bm = bm0; bm.c_off += self ? T.send_off : T.recv_off;
      box_copy<C>(st, w1, self ? pk : w0, bm, BOX_C2A, false, launches);
    }
  }
}

// backward: dst-side array `arr` -> (pack, recv-side chunks) pk -> (exchange) arr's buffer (arr is dead once the exchange
// starts) -> (unpack) src-side array `out` (any buffer but pk and arr; w is kept in the signature for the callers'
// ping-pong).  The rank's own chunk goes arr -> out in one strided copy before the exchange overwrites arr.
template <class C>
inline void run_stage_backward(const Stage &S, const Mesh &M, C *arr, C *w, C *out, C *pk, cudaStream_t st, long long *launches) {
  (void)w;
  for (const auto &T : S.tr) {
    if (!T.self_maps.empty()) {
      for (const auto &bm : T.self_maps) box_copy<C>(st, arr, out, bm, BOX_A2C, T.send_sign, launches);
      continue;
    }
    for (const auto &bm0 : T.recv_maps) {
      BoxMap bm = bm0; bm.c_off += T.recv_off;
      box_copy<C>(st, arr, pk, bm, BOX_A2C, false, launches);
    }
  }
  exchange_chunks<C>(S, M, pk, arr, false, st, false);
  for (const auto &T : S.tr) {
    if (!T.self_maps.empty()) continue;
    const bool self = T.peer == M.rank;
    for (const auto &bm0 : T.send_maps) {
      BoxMap bm = bm0; bm.c_off += self ? T.recv_off : T.send_off;
      box_copy<C>(st, out, self ? pk : arr, bm, BOX_C2A, T.send_sign, launches);
    }
  }
}

// -----------------------------------------------------------------------------------------------
// cuFFT plumbing
// -----------------------------------------------------------------------------------------------
template <class R> struct FftType;
template <> struct FftType<double> {
  static constexpr cufftType c2c = CUFFT_Z2Z, c2r = CUFFT_Z2D, r2c = CUFFT_D2Z;
  static void exec_c2c(cufftHandle h, double2 *d, int dir) { PNB_CUFFT(cufftExecZ2Z(h, d, d, dir)); }
  static void exec_c2r(cufftHandle h, double2 *in, double *out) { PNB_CUFFT(cufftExecZ2D(h, in, out)); }
  static void exec_r2c(cufftHandle h, double *in, double2 *out) { PNB_CUFFT(cufftExecD2Z(h, in, out)); }
};
template <> struct FftType<float> {
  static constexpr cufftType c2c = CUFFT_C2C, c2r = CUFFT_C2R, r2c = CUFFT_R2C;
  static void exec_c2c(cufftHandle h, float2 *d, int dir) { PNB_CUFFT(cufftExecC2C(h, d, d, dir)); }
  static void exec_c2r(cufftHandle h, float2 *in, float *out) { PNB_CUFFT(cufftExecC2R(h, in, out)); }
  static void exec_r2c(cufftHandle h, float *in, float2 *out) { PNB_CUFFT(cufftExecR2C(h, in, out)); }
};

inline cufftHandle make_plan_1d(long long n, long long stride, long long dist, long long batch, cufftType type,
                                long long odist_override, cudaStream_t st) {
  cufftHandle h = 0;
  if (batch <= 0 || n <= 0) return 0;
  PNB_CUFFT(cufftCreate(&h));
  long long nn[1] = {n};
  long long inembed[1] = {n}, onembed[1] = {n};
  size_t ws = 0;
  long long idist = dist, odist = dist;
  if (type == CUFFT_Z2D || type == CUFFT_C2R) { idist = n / 2 + 1; odist = odist_override; inembed[0] = n / 2 + 1; }
  if (type == CUFFT_D2Z || type == CUFFT_R2C) { idist = odist_override; odist = n / 2 + 1; onembed[0] = n / 2 + 1; }
  PNB_CUFFT(cufftMakePlanMany64(h, 1, nn, inembed, stride, idist, onembed, stride, odist, type, batch, &ws));
  PNB_CUFFT(cufftSetStream(h, st));
  return h;
}

}  // namespace pnb
