// hostcomm.cpp -- host control plane: process bootstrap + tiny scoped all-gathers over TCP.
//
// Replaces what the reference gets from MPI_Init / MPI_Comm_split / MPI_Allreduce for *setup-sized*
// messages only (BLACS/SRC/blacs_pinfo_.c:14-27, blacs_map_.c:106-118, igamn2d_.c:276-289): rank
// discovery, NCCL unique-id exchange, PCHK1MAT/PCHK2MAT argument consistency and BLACS_BARRIER.
// Matrix data never travels here -- that is NCCL over NVLink (ncclw.cpp, lu_dist.cu).
//
// Topology: a star.  Rank 0 runs a server thread; every rank (0 included) holds one client socket.
// A collective = each member sends {group key, nmembers, index, payload}; when the server has all
// nmembers payloads of a key it answers each member with the concatenation in index order.
#include "common.h"

#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdarg>
#include <map>
#include <mutex>
#include <thread>

namespace slb {

void fatal(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    fprintf(stderr, "[scalapack_b200] FATAL: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    fflush(stderr);
    abort();
}

int verbose()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("SLB200_VERBOSE"); v = e ? atoi(e) : 0; }
    return v > (int)opt("verbose", 0) ? v : (int)opt("verbose", 0);
}

void vlog(int level, const char *fmt, ...)
{
    if (verbose() < level) return;
    va_list ap; va_start(ap, fmt);
    fprintf(stderr, "[scalapack_b200 r%d] ", hc_rank());
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
}

struct ReqHeader { uint64_t group; int32_t nmembers; int32_t index; uint64_t len; };

struct HostComm {
    int rank = 0, size = 1;
    int fd = -1;                 // client socket to the server
    std::thread server;
    int listen_fd = -1;
    bool up = false;
};

static HostComm g_hc;
static std::mutex g_mu;

static void write_full(int fd, const void *buf, size_t n)
{
    const char *p = (const char *)buf;
    while (n) {
        ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
        if (w < 0) { if (errno == EINTR) continue; fatal("hostcomm send failed: %s", strerror(errno)); }
        p += w; n -= (size_t)w;
    }
}
static bool read_full(int fd, void *buf, size_t n)
{
    char *p = (char *)buf;
    while (n) {
        ssize_t r = ::recv(fd, p, n, 0);
        if (r == 0) return false;
        if (r < 0) { if (errno == EINTR) continue; return false; }
        p += r; n -= (size_t)r;
    }
    return true;
}

static int env_int(const char *a, const char *b, const char *c, int dflt)
{
    const char *names[3] = { a, b, c };
    for (const char *n : names) { if (!n) continue; const char *e = getenv(n); if (e && *e) return atoi(e); }
    return dflt;
}

static void server_loop(int listen_fd, int size)
{
    std::vector<int> fds;
    for (int i = 0; i < size; ++i) {
        int c = ::accept(listen_fd, nullptr, nullptr);
        if (c < 0) { if (errno == EINTR) { --i; continue; } fatal("hostcomm accept failed: %s", strerror(errno)); }
        int one = 1; setsockopt(c, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
        int32_t r; if (!read_full(c, &r, sizeof(r))) fatal("hostcomm: handshake failed");
        if (r < 0 || r >= size) { ::close(c); --i; continue; }      // not one of this job's ranks: drop the connection
        fds.push_back(c);
    }
    ::close(listen_fd);
    struct Pending { int n = 0; size_t len = 0; std::vector<std::string> parts; std::vector<int> who; };
    std::map<uint64_t, Pending> pend;
    std::vector<pollfd> pfds(fds.size());
    size_t alive = fds.size();
    while (alive > 0) {
        for (size_t i = 0; i < fds.size(); ++i) { pfds[i].fd = fds[i]; pfds[i].events = POLLIN; pfds[i].revents = 0; }
        int pr = ::poll(pfds.data(), pfds.size(), -1);
        if (pr < 0) { if (errno == EINTR) continue; break; }
        for (size_t i = 0; i < fds.size(); ++i) {
            if (fds[i] < 0 || !(pfds[i].revents & (POLLIN | POLLHUP | POLLERR))) continue;
            ReqHeader h;
            if (!read_full(fds[i], &h, sizeof(h))) { ::close(fds[i]); fds[i] = -1; --alive; continue; }
            // bound what the wire may ask for: collectives of this control plane carry descriptors, ids and small vectors
            if (h.len > ((uint64_t)256 << 20) || h.nmembers < 1 || h.nmembers > size) fatal("hostcomm: malformed request (len %llu, members %d)", (unsigned long long)h.len, h.nmembers);
            std::string payload(h.len, '\0');
            if (h.len && !read_full(fds[i], &payload[0], h.len)) { ::close(fds[i]); fds[i] = -1; --alive; continue; }
            Pending &p = pend[h.group];
            if (p.parts.empty()) { p.parts.resize(h.nmembers); p.who.assign(h.nmembers, -1); p.len = h.len; }
            if (h.index < 0 || h.index >= (int)p.parts.size() || p.len != h.len || p.who[h.index] != -1)
                fatal("hostcomm: inconsistent collective (group %llx index %d)", (unsigned long long)h.group, h.index);
            p.parts[h.index] = std::move(payload); p.who[h.index] = fds[i]; ++p.n;
            if (p.n == (int)p.parts.size()) {
                std::string all; all.reserve(p.len * p.parts.size());
                for (auto &s : p.parts) all += s;
                for (int fd : p.who) write_full(fd, all.data(), all.size());
                pend.erase(h.group);
            }
        }
    }
}

static void bootstrap()
{
    if (g_hc.up) return;
    g_hc.rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", 0);
    g_hc.size = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", 1);
    if (g_hc.size < 1) g_hc.size = 1;
    if (g_hc.rank < 0 || g_hc.rank >= g_hc.size) fatal("bad RANK=%d for WORLD_SIZE=%d", g_hc.rank, g_hc.size);
    g_hc.up = true;
    if (g_hc.size == 1) return;
    const char *addr = getenv("MASTER_ADDR"); if (!addr || !*addr) addr = "127.0.0.1";
    int port = env_int("SLB200_PORT", nullptr, nullptr, 0);
    if (port == 0) port = env_int("MASTER_PORT", nullptr, nullptr, 29500) + env_int("SLB200_PORT_OFFSET", nullptr, nullptr, 23);
    if (g_hc.rank == 0) {
        int lf = ::socket(AF_INET, SOCK_STREAM, 0);
        int one = 1; setsockopt(lf, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
        // single-node jobs (MASTER_ADDR is loopback or unset) listen on loopback only; SLB200_BIND_ANY=1 restores INADDR_ANY
        sockaddr_in sa; memset(&sa, 0, sizeof(sa)); sa.sin_family = AF_INET;
        const bool loop = strncmp(addr, "127.", 4) == 0 || strcmp(addr, "localhost") == 0;
        sa.sin_addr.s_addr = htonl((loop && env_int("SLB200_BIND_ANY", nullptr, nullptr, 0) == 0) ? INADDR_LOOPBACK : INADDR_ANY);
        sa.sin_port = htons((uint16_t)port);
        if (::bind(lf, (sockaddr *)&sa, sizeof(sa)) != 0) fatal("hostcomm: cannot bind port %d: %s", port, strerror(errno));
        if (::listen(lf, g_hc.size + 8) != 0) fatal("hostcomm: listen failed: %s", strerror(errno));
        g_hc.listen_fd = lf;
        g_hc.server = std::thread(server_loop, lf, g_hc.size);
        g_hc.server.detach();
    }
    // client side (rank 0 connects to itself)
    addrinfo hints; memset(&hints, 0, sizeof(hints)); hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
    addrinfo *res = nullptr;
    char ports[16]; snprintf(ports, sizeof(ports), "%d", port);
    const char *target = g_hc.rank == 0 ? "127.0.0.1" : addr;
    if (getaddrinfo(target, ports, &hints, &res) != 0 || !res) fatal("hostcomm: cannot resolve %s", target);
    int fd = -1;
    auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        fd = ::socket(AF_INET, SOCK_STREAM, 0);
        if (::connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
        ::close(fd); fd = -1;
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(300)) fatal("hostcomm: timeout connecting to %s:%d", target, port);
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    freeaddrinfo(res);
    int one = 1; setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    int32_t r = g_hc.rank; write_full(fd, &r, sizeof(r));
    g_hc.fd = fd;
}

HostComm *hostcomm() { std::lock_guard<std::mutex> lk(g_mu); bootstrap(); return &g_hc; }
int hc_rank() { return hostcomm()->rank; }
int hc_size() { return hostcomm()->size; }

void hc_allgather(uint64_t group, int nmembers, int index, const void *in, void *out, size_t len)
{
    HostComm *hc = hostcomm();
    if (nmembers <= 1) { if (len) memcpy(out, in, len); return; }
    if (hc->size == 1) fatal("hostcomm: collective over %d members in a single-process run", nmembers);
    std::lock_guard<std::mutex> lk(g_mu);
    ReqHeader h{ group, nmembers, index, (uint64_t)len };
    write_full(hc->fd, &h, sizeof(h));
    if (len) write_full(hc->fd, in, len);
    if (!read_full(hc->fd, out, len * (size_t)nmembers)) fatal("hostcomm: peer closed during collective");
}

void hc_shutdown()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_hc.fd >= 0) { ::close(g_hc.fd); g_hc.fd = -1; }
    g_hc.up = false;
}

}  // namespace slb
