// Sys.cpp -- process bootstrap (replaces MPI_Init/PetscInitialize/Zoltan_Initialize of src/Sys/Sys.cpp:31-63).
#include "Sys.h"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>

namespace pacmensl {

static std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), ::tolower);
  return s;
}

PartitioningType str2part(std::string str) {
  str = lower(str);
  if (str == "graph") return PartitioningType::GRAPH;
  if (str == "hypergraph") return PartitioningType::HYPERGRAPH;
  if (str == "hier" || str == "hierarchical") return PartitioningType::HIERARCHICAL;
  return PartitioningType::BLOCK;
}
std::string part2str(PartitioningType part) {
  switch (part) {
    case PartitioningType::GRAPH: return "Graph";
    case PartitioningType::HYPERGRAPH: return "Hypergraph";
    case PartitioningType::HIERARCHICAL: return "Hierarchical";
    default: return "Block";
  }
}
PartitioningApproach str2partapproach(std::string str) {
  str = lower(str);
  if (str == "from_scratch" || str == "partition" || str == "fromscratch") return PartitioningApproach::FROMSCRATCH;
  if (str == "refine") return PartitioningApproach::REFINE;
  return PartitioningApproach::REPARTITION;
}
std::string partapproach2str(PartitioningApproach a) {
  switch (a) {
    case PartitioningApproach::FROMSCRATCH: return "FromScratch";
    case PartitioningApproach::REFINE: return "Refine";
    default: return "Repart";
  }
}

double round2digit(double x) {
  if (x == 0.0) return x;
  double p1 = std::pow(10.0e0, std::round(std::log10(x) - std::sqrt(0.1e0)) - 1.0e0);
  return std::trunc(x / p1 + 0.55e0) * p1;
}

// NCCL unique-id rendezvous for stand-alone C++ programs launched with torchrun-style environment
// variables (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT): rank 0 serves the id over TCP.
static int exchange_id(char id[FSPCOMM_ID_BYTES], int rank, int size) {
  const char *addr = std::getenv("MASTER_ADDR");
  const char *port_s = std::getenv("MASTER_PORT");
  int         port = (port_s ? std::atoi(port_s) : 29500) + 17;
  if (rank == 0) {
    if (fspcomm_unique_id(id)) return -1;
    int srv = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa{};
    sa.sin_family = AF_INET;
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    sa.sin_port = htons((uint16_t) port);
    if (bind(srv, (sockaddr *) &sa, sizeof(sa)) != 0 || listen(srv, size) != 0) { close(srv); return -1; }
    for (int r = 1; r < size; ++r) {
      int c = accept(srv, nullptr, nullptr);
      if (c < 0) { close(srv); return -1; }
      ssize_t w = write(c, id, FSPCOMM_ID_BYTES);
      close(c);
      if (w != FSPCOMM_ID_BYTES) { close(srv); return -1; }
    }
    close(srv);
    return 0;
  }
  for (int attempt = 0; attempt < 600; ++attempt) {
    int         c = socket(AF_INET, SOCK_STREAM, 0);
    sockaddr_in sa{};
    sa.sin_family = AF_INET;
    sa.sin_port = htons((uint16_t) port);
    inet_pton(AF_INET, addr ? addr : "127.0.0.1", &sa.sin_addr);
    if (connect(c, (sockaddr *) &sa, sizeof(sa)) == 0) {
      size_t got = 0;
      while (got < FSPCOMM_ID_BYTES) {
        ssize_t r = read(c, id + got, FSPCOMM_ID_BYTES - got);
        if (r <= 0) break;
        got += (size_t) r;
      }
      close(c);
      if (got == FSPCOMM_ID_BYTES) return 0;
    } else {
      close(c);
    }
    usleep(100000);
  }
  return -1;
}

static bool g_initialized = false;

int PACMENSLInit(int *, char ***, const char *) {
  if (g_initialized) return 0;
  const char *r = std::getenv("RANK"), *w = std::getenv("WORLD_SIZE"), *lr = std::getenv("LOCAL_RANK");
  int rank = r ? std::atoi(r) : 0, size = w ? std::atoi(w) : 1, local = lr ? std::atoi(lr) : 0;
  int ndev = 0;
  if (fsp_device_count(&ndev) || ndev <= 0) {
    printf("PACMENSL (B200): no CUDA device visible; this library has no CPU path.\n");
    return -1;
  }
  if (fsp_device_set(local % ndev)) return -1;
  if (size > 1) {
    char id[FSPCOMM_ID_BYTES];
    if (exchange_id(id, rank, size)) return -1;
    if (pacmensl_comm_world_init(id, rank, size)) return -1;
  }
  g_initialized = true;
  return 0;
}

int PACMENSLFinalize() {
  pacmensl_comm_world_finalize();
  g_initialized = false;
  return 0;
}

Environment::Environment() {
  if (PACMENSLInit(nullptr, nullptr, nullptr)) throw std::runtime_error("PACMENSLInit failed");
  initialized = true;
}
Environment::Environment(int *argc, char ***argv, const char *help) {
  if (PACMENSLInit(argc, argv, help)) throw std::runtime_error("PACMENSLInit failed");
  initialized = true;
}
Environment::~Environment() {
  if (initialized) PACMENSLFinalize();
}

}  // namespace pacmensl
