// pnfft-b200 mini-MPI (see include/mpi.h): the host control plane of a one-process-per-GPU job.
//
// PNFFT only ever uses communicators that span all ranks of the job (MPI_COMM_WORLD, its
// duplicates and Cartesian meshes over it: reference util/util.c:23-40,
// kernel/ndft-parallel.c:899-908), and it only needs rank/size/mesh queries plus a few tiny
// host collectives (timer max-reduce kernel/timer.c:79-86, test drivers' error norms).  Those are
// served through a TCP star rooted at rank 0.  Bulk data never travels here; it goes over NCCL.
#include <mpi.h>

#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Comm {
  bool used = false;
  int ndims = 0;           // 0: not Cartesian
  int dims[3] = {1, 1, 1};
  int periods[3] = {1, 1, 1};
};

struct World {
  bool init = false;
  int rank = 0, size = 1;
  int listen_fd = -1;
  std::vector<int> peer;   // rank 0: socket of every rank; others: peer[0] = socket to rank 0
  std::vector<Comm> comms; // index = handle
} W;

[[noreturn]] void die(const char *msg) {
  fprintf(stderr, "pnfft-b200 mini-MPI (rank %d): %s (%s)\n", W.rank, msg, strerror(errno));
  abort();
}

double now_s() {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

int env_int(const char *const *names, int dflt) {
  for (int i = 0; names[i]; i++) {
    const char *v = getenv(names[i]);
    if (v && *v) return atoi(v);
  }
  return dflt;
}

void send_all(int fd, const void *buf, size_t n) {
  const char *p = (const char *)buf;
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) { if (errno == EINTR) continue; die("send failed"); }
    p += k; n -= (size_t)k;
  }
}
void recv_all(int fd, void *buf, size_t n) {
  char *p = (char *)buf;
  while (n) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; die("recv failed (peer gone?)"); }
    p += k; n -= (size_t)k;
  }
}

size_t type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
    case MPI_LONG: case MPI_DOUBLE: return 8;
    case MPI_LONG_DOUBLE: return sizeof(long double);
  }
  die("unknown datatype");
}

template <class T> void combine_t(T *acc, const T *in, int count, MPI_Op op) {
  for (int i = 0; i < count; i++) {
    if (op == MPI_SUM) acc[i] += in[i];
    else if (op == MPI_MAX) { if (in[i] > acc[i]) acc[i] = in[i]; }
    else if (op == MPI_MIN) { if (in[i] < acc[i]) acc[i] = in[i]; }
  }
}
void combine(void *acc, const void *in, int count, MPI_Datatype t, MPI_Op op) {
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: combine_t((unsigned char *)acc, (const unsigned char *)in, count, op); break;
    case MPI_INT: combine_t((int *)acc, (const int *)in, count, op); break;
    case MPI_UNSIGNED: combine_t((unsigned *)acc, (const unsigned *)in, count, op); break;
    case MPI_LONG: combine_t((long *)acc, (const long *)in, count, op); break;
    case MPI_FLOAT: combine_t((float *)acc, (const float *)in, count, op); break;
    case MPI_DOUBLE: combine_t((double *)acc, (const double *)in, count, op); break;
    case MPI_LONG_DOUBLE: combine_t((long double *)acc, (const long double *)in, count, op); break;
    default: die("unknown datatype");
  }
}

// Rank 0 listens on MASTER_ADDR (loopback by default), never on every interface; a peer has to present the job's
// token (PNFFT_B200_TOKEN, else derived from the launcher's run id / port / world size) before its rank is believed, and
// neither accept nor the handshake waits for ever.
uint64_t job_token() {
  const char *src[] = {getenv("PNFFT_B200_TOKEN"), getenv("TORCHELASTIC_RUN_ID"), getenv("MASTER_PORT"), getenv("WORLD_SIZE")};
  uint64_t h = 1469598103934665603ull;     // FNV-1a
  for (const char *s : src) {
    if (!s) s = "-";
    for (; *s; s++) { h ^= (unsigned char)*s; h *= 1099511628211ull; }
    h ^= 0xff; h *= 1099511628211ull;
  }
  return h;
}
struct Hello { uint32_t magic; int32_t rank; uint64_t token; };
constexpr uint32_t kMagic = 0x32424e50u;   // "PNB2"

bool recv_timed(int fd, void *buf, size_t n, int timeout_ms) {
  char *p = (char *)buf;
  while (n) {
    pollfd pf{fd, POLLIN, 0};
    const int r = poll(&pf, 1, timeout_ms);
    if (r == 0) return false;
    if (r < 0) { if (errno == EINTR) continue; return false; }
    ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; return false; }
    p += k; n -= (size_t)k;
  }
  return true;
}

void rendezvous() {
  static const char *port_names[] = {"MASTER_PORT", nullptr};
  static const char *off_names[] = {"PNFFT_B200_PORT_OFFSET", nullptr};
  static const char *to_names[] = {"PNFFT_B200_RENDEZVOUS_TIMEOUT_S", nullptr};
  const int port = env_int(port_names, 29500) + env_int(off_names, 1007);
  const int timeout_s = env_int(to_names, 300);
  const char *addr = getenv("MASTER_ADDR");
  if (!addr || !*addr) addr = "127.0.0.1";
  const int one = 1;
  const uint64_t token = job_token();
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
  char ps[16]; snprintf(ps, sizeof ps, "%d", port);
  if (getaddrinfo(addr, ps, &hints, &res) != 0 || !res) die("cannot resolve MASTER_ADDR");
  if (W.rank == 0) {
    W.listen_fd = socket(AF_INET, SOCK_STREAM, 0);
    if (W.listen_fd < 0) die("socket");
    setsockopt(W.listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    if (bind(W.listen_fd, res->ai_addr, res->ai_addrlen) < 0) {
      // MASTER_ADDR is not an address of this host (a name that resolves elsewhere): loopback is the safe default
      sockaddr_in sa{};
      sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_LOOPBACK); sa.sin_port = htons((uint16_t)port);
      if (bind(W.listen_fd, (sockaddr *)&sa, sizeof sa) < 0) die("bind (set PNFFT_B200_PORT_OFFSET to move the rendezvous port)");
    }
    freeaddrinfo(res);
    if (listen(W.listen_fd, W.size + 8) < 0) die("listen");
    W.peer.assign((size_t)W.size, -1);
    const double t_end = now_s() + timeout_s;
    for (int have = 1; have < W.size;) {
      pollfd pf{W.listen_fd, POLLIN, 0};
      const double left = t_end - now_s();
      if (left <= 0) die("rendezvous timed out waiting for the other ranks (PNFFT_B200_RENDEZVOUS_TIMEOUT_S)");
      const int pr = poll(&pf, 1, (int)(left * 1000) + 1);
      if (pr < 0 && errno == EINTR) continue;
      if (pr <= 0) continue;
      int fd = accept(W.listen_fd, nullptr, nullptr);
      if (fd < 0) continue;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      Hello h{};
      // a stray or foreign connection is dropped, it cannot claim a rank or stall the job
      if (!recv_timed(fd, &h, sizeof h, 10000) || h.magic != kMagic || h.token != token || h.rank <= 0 || h.rank >= W.size ||
          W.peer[(size_t)h.rank] != -1) { close(fd); continue; }
      W.peer[(size_t)h.rank] = fd;
      have++;
    }
  } else {
    int fd = -1;
    for (int attempt = 0; attempt < timeout_s * 10; attempt++) {
      fd = socket(AF_INET, SOCK_STREAM, 0);
      if (fd < 0) die("socket");
      if (connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
      close(fd); fd = -1;
      usleep(100000);
    }
    freeaddrinfo(res);
    if (fd < 0) die("cannot reach rank 0");
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    Hello h{kMagic, (int32_t)W.rank, token};
    send_all(fd, &h, sizeof h);
    W.peer.assign(1, fd);
  }
}

void ensure_init() {
  if (W.init) return;
  static const char *rank_names[] = {"PNFFT_B200_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", nullptr};
  static const char *size_names[] = {"PNFFT_B200_WORLD_SIZE", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", nullptr};
  W.rank = env_int(rank_names, 0);
  W.size = env_int(size_names, 1);
  if (W.size < 1 || W.rank < 0 || W.rank >= W.size) { W.rank = 0; W.size = 1; }
  W.comms.assign(3, Comm());
  W.comms[MPI_COMM_WORLD].used = true;
  W.comms[MPI_COMM_SELF].used = true;
  W.init = true;
  if (W.size > 1) rendezvous();
}

Comm *get(MPI_Comm c) {
  ensure_init();
  if (c <= 0 || (size_t)c >= W.comms.size() || !W.comms[(size_t)c].used) return nullptr;
  return &W.comms[(size_t)c];
}

MPI_Comm new_comm(const Comm &src) {
  for (size_t i = 3; i < W.comms.size(); i++)
    if (!W.comms[i].used) { W.comms[i] = src; W.comms[i].used = true; return (MPI_Comm)i; }
  W.comms.push_back(src);
  W.comms.back().used = true;
  return (MPI_Comm)(W.comms.size() - 1);
}

inline bool is_self(MPI_Comm c) { return c == MPI_COMM_SELF || W.size == 1; }

}  // namespace

extern "C" {

int MPI_Init(int *, char ***) { ensure_init(); return MPI_SUCCESS; }
int MPI_Initialized(int *flag) { *flag = W.init ? 1 : 0; return MPI_SUCCESS; }

int MPI_Finalize(void) {
  if (!W.init) return MPI_SUCCESS;
  if (W.size > 1) MPI_Barrier(MPI_COMM_WORLD);
  for (int fd : W.peer) if (fd >= 0) close(fd);
  if (W.listen_fd >= 0) close(W.listen_fd);
  W = World();
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int errorcode) { fprintf(stderr, "MPI_Abort(%d)\n", errorcode); abort(); }

int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  if (!get(comm)) return MPI_ERR_OTHER;
  *rank = comm == MPI_COMM_SELF ? 0 : W.rank;
  return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm comm, int *size) {
  if (!get(comm)) return MPI_ERR_OTHER;
  *size = comm == MPI_COMM_SELF ? 1 : W.size;
  return MPI_SUCCESS;
}

int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm) {
  Comm *c = get(comm);
  if (!c || comm == MPI_COMM_SELF) { *newcomm = comm == MPI_COMM_SELF ? MPI_COMM_SELF : MPI_COMM_NULL; return c ? MPI_SUCCESS : MPI_ERR_OTHER; }
  *newcomm = new_comm(*c);
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm *comm) {
  if (*comm >= 3 && get(*comm)) W.comms[(size_t)*comm].used = false;
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int, MPI_Comm *comm_cart) {
  Comm *c = get(comm);
  *comm_cart = MPI_COMM_NULL;
  if (!c || ndims < 1 || ndims > 3) return MPI_ERR_OTHER;
  int size, prod = 1;
  MPI_Comm_size(comm, &size);
  for (int t = 0; t < ndims; t++) prod *= dims[t];
  if (prod != size) return MPI_ERR_OTHER;
  Comm k;
  k.ndims = ndims;
  for (int t = 0; t < ndims; t++) { k.dims[t] = dims[t]; k.periods[t] = periods ? periods[t] : 1; }
  *comm_cart = new_comm(k);
  return MPI_SUCCESS;
}
int MPI_Cartdim_get(MPI_Comm comm, int *ndims) {
  Comm *c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  *ndims = c->ndims;
  return MPI_SUCCESS;
}
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords) {
  Comm *c = get(comm);
  if (!c || c->ndims == 0) return MPI_ERR_OTHER;
  // row-major rank order, last mesh dimension varies fastest (as every MPI does)
  for (int t = c->ndims - 1; t >= 0; t--) {
    if (t < maxdims) coords[t] = rank % c->dims[t];
    rank /= c->dims[t];
  }
  return MPI_SUCCESS;
}
int MPI_Cart_rank(MPI_Comm comm, const int *coords, int *rank) {
  Comm *c = get(comm);
  if (!c || c->ndims == 0) return MPI_ERR_OTHER;
  int r = 0;
  for (int t = 0; t < c->ndims; t++) {
    int q = coords[t] % c->dims[t];
    if (q < 0) q += c->dims[t];
    r = r * c->dims[t] + q;
  }
  *rank = r;
  return MPI_SUCCESS;
}
int MPI_Cart_get(MPI_Comm comm, int maxdims, int *dims, int *periods, int *coords) {
  Comm *c = get(comm);
  if (!c || c->ndims == 0) return MPI_ERR_OTHER;
  int me;
  MPI_Comm_rank(comm, &me);
  for (int t = 0; t < c->ndims && t < maxdims; t++) { dims[t] = c->dims[t]; periods[t] = c->periods[t]; }
  return MPI_Cart_coords(comm, me, maxdims, coords);
}

int MPI_Barrier(MPI_Comm comm) {
  if (!get(comm)) return MPI_ERR_OTHER;
  if (is_self(comm)) return MPI_SUCCESS;
  char b = 1;
  if (W.rank == 0) {
    for (int r = 1; r < W.size; r++) recv_all(W.peer[(size_t)r], &b, 1);
    for (int r = 1; r < W.size; r++) send_all(W.peer[(size_t)r], &b, 1);
  } else {
    send_all(W.peer[0], &b, 1);
    recv_all(W.peer[0], &b, 1);
  }
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
  if (!get(comm)) return MPI_ERR_OTHER;
  if (is_self(comm)) return MPI_SUCCESS;
  const size_t bytes = (size_t)count * type_size(type);
  if (W.rank == 0) {
    if (root != 0) recv_all(W.peer[(size_t)root], buf, bytes);
    for (int r = 1; r < W.size; r++) if (r != root) send_all(W.peer[(size_t)r], buf, bytes);
  } else if (W.rank == root) {
    send_all(W.peer[0], buf, bytes);
  } else {
    recv_all(W.peer[0], buf, bytes);
  }
  return MPI_SUCCESS;
}

static int reduce_impl(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, bool all, MPI_Comm comm) {
  if (!get(comm)) return MPI_ERR_OTHER;
  const size_t bytes = (size_t)count * type_size(type);
  if (is_self(comm)) { if (recvbuf != sendbuf) memcpy(recvbuf, sendbuf, bytes); return MPI_SUCCESS; }
  if (W.rank == 0) {
    std::vector<char> acc(bytes), in(bytes);
    memcpy(acc.data(), sendbuf, bytes);
    for (int r = 1; r < W.size; r++) {          // fixed rank order: deterministic sums
      recv_all(W.peer[(size_t)r], in.data(), bytes);
      combine(acc.data(), in.data(), count, type, op);
    }
    if (all) { for (int r = 1; r < W.size; r++) send_all(W.peer[(size_t)r], acc.data(), bytes); memcpy(recvbuf, acc.data(), bytes); }
    else if (root == 0) memcpy(recvbuf, acc.data(), bytes);
    else send_all(W.peer[(size_t)root], acc.data(), bytes);
  } else {
    send_all(W.peer[0], sendbuf, bytes);
    if (all || W.rank == root) recv_all(W.peer[0], recvbuf, bytes);
  }
  return MPI_SUCCESS;
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm) {
  return reduce_impl(sendbuf, recvbuf, count, type, op, root, false, comm);
}
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  return reduce_impl(sendbuf, recvbuf, count, type, op, 0, true, comm);
}

double MPI_Wtime(void) {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

}  // extern "C"
