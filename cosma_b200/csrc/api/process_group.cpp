// cosma::pg over TCP (see include/cosma/process_group.hpp). Control plane only: rank/size, the ncclUniqueId broadcast,
// barriers and the small gathers of tests and miniapps. Matrix data never travels here.
#include <cosma/process_group.hpp>

#include <algorithm>
#include <arpa/inet.h>
#include <cerrno>
#include <chrono>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <stdexcept>
#include <string>
#include <sys/socket.h>
#include <sys/time.h>
#include <thread>
#include <unistd.h>

namespace cosma {
namespace pg {

struct group {
    std::vector<int> members;  // world ranks, in group order
    int my_pos = -1;
    std::uint64_t gid = 0;
    std::uint64_t children = 0;  // splits / dups made from this group (same on every member)
};

namespace {

struct frame_header {
    std::uint64_t gid;
    std::int32_t tag;
    std::int32_t src;  // world rank of the sender
    std::uint64_t bytes;
};
struct message {
    frame_header h;
    std::vector<char> data;
};

struct state_t {
    bool up = false;
    int world_rank = 0, world_size = 1, local_rank = 0;
    std::vector<int> sock;                    // by world rank; -1 for self
    std::vector<std::deque<message>> parked;  // frames that arrived before somebody asked for them, by world rank
    group* world = nullptr;
    std::recursive_mutex mu;
};
state_t& S() {
    static state_t s;
    return s;
}

constexpr std::int32_t TAG_COLL = -77;  // internal collectives

[[noreturn]] void fail(const std::string& what) { throw std::runtime_error("cosma::pg: " + what); }

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

void write_all(int fd, const void* buf, std::size_t n) {
    const char* p = static_cast<const char*>(buf);
    while (n > 0) {
        const ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
        if (w < 0) {
            if (errno == EINTR) continue;
            fail(std::string("send: ") + std::strerror(errno));
        }
        p += w;
        n -= static_cast<std::size_t>(w);
    }
}
void read_all(int fd, void* buf, std::size_t n) {
    char* p = static_cast<char*>(buf);
    while (n > 0) {
        const ssize_t r = ::recv(fd, p, n, 0);
        if (r == 0) fail("peer closed the connection");
        if (r < 0) {
            if (errno == EINTR) continue;
            if (errno == EAGAIN || errno == EWOULDBLOCK) fail("no message within COSMA_B200_PG_RECV_TIMEOUT seconds: giving up");
            fail(std::string("recv: ") + std::strerror(errno));
        }
        p += r;
        n -= static_cast<std::size_t>(r);
    }
}

void tune(int fd) {
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    // COSMA_B200_PG_RECV_TIMEOUT [s] (unset: wait forever, like MPI): a rank that hears nothing for that long gives up instead of
    // outliving a killed launcher -- test launchers set it so that a protocol error can never leave processes parked on the GPUs
    const int limit = env_int("COSMA_B200_PG_RECV_TIMEOUT", 0);
    if (limit > 0) {
        timeval tv{};
        tv.tv_sec = limit;
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    }
}

int listen_on(int port, int* bound_port) {
    const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) fail("socket()");
    int one = 1;
    setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_ANY);
    a.sin_port = htons(static_cast<uint16_t>(port));
    if (::bind(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0) fail("bind to port " + std::to_string(port) + ": " + std::strerror(errno));
    if (::listen(fd, 128) != 0) fail("listen()");
    socklen_t len = sizeof(a);
    getsockname(fd, reinterpret_cast<sockaddr*>(&a), &len);
    if (bound_port) *bound_port = ntohs(a.sin_port);
    return fd;
}

int connect_to(std::uint32_t addr_be, int port, double timeout_s) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
        if (fd < 0) fail("socket()");
        sockaddr_in a{};
        a.sin_family = AF_INET;
        a.sin_addr.s_addr = addr_be;
        a.sin_port = htons(static_cast<uint16_t>(port));
        if (::connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) == 0) {
            tune(fd);
            return fd;
        }
        ::close(fd);
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
            fail("could not connect to port " + std::to_string(port) + " within " + std::to_string(static_cast<int>(timeout_s)) + " s");
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
}

std::uint32_t resolve(const std::string& host) {
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host.c_str(), nullptr, &hints, &res) != 0 || !res) fail("cannot resolve MASTER_ADDR '" + host + "'");
    const std::uint32_t a = reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr.s_addr;
    freeaddrinfo(res);
    return a;
}

// full mesh: every pair of ranks keeps one socket
void connect_world() {
    state_t& s = S();
    const int R = s.world_rank, N = s.world_size;
    s.sock.assign(N, -1);
    s.parked.assign(N, {});
    if (N == 1) return;
    const char* host = std::getenv("MASTER_ADDR");
    const std::uint32_t master = resolve(host && *host ? host : "127.0.0.1");
    const int base_port = env_int("COSMA_B200_PG_PORT", env_int("MASTER_PORT", 29500) + 1);
    const double timeout = env_int("COSMA_B200_PG_TIMEOUT", 120);
    struct entry { std::uint32_t addr; std::int32_t port; };
    std::vector<entry> table(N);
    int my_listen = -1, my_port = 0;
    if (R == 0) {
        my_listen = listen_on(base_port, &my_port);
        table[0] = {master, my_port};
        for (int i = 1; i < N; ++i) {
            sockaddr_in pa{};
            socklen_t len = sizeof(pa);
            const int fd = ::accept(my_listen, reinterpret_cast<sockaddr*>(&pa), &len);
            if (fd < 0) fail("accept()");
            tune(fd);
            std::int32_t hello[2];
            read_all(fd, hello, sizeof(hello));
            if (hello[0] <= 0 || hello[0] >= N || s.sock[hello[0]] != -1) fail("unexpected rank in rendezvous");
            s.sock[hello[0]] = fd;
            table[hello[0]] = {pa.sin_addr.s_addr, hello[1]};
        }
        for (int i = 1; i < N; ++i) write_all(s.sock[i], table.data(), sizeof(entry) * N);
        ::close(my_listen);
    } else {
        my_listen = listen_on(0, &my_port);
        s.sock[0] = connect_to(master, base_port, timeout);
        const std::int32_t hello[2] = {R, my_port};
        write_all(s.sock[0], hello, sizeof(hello));
        read_all(s.sock[0], table.data(), sizeof(entry) * N);
        // connect to the lower ranks (>0), accept the higher ones
        for (int j = 1; j < R; ++j) {
            s.sock[j] = connect_to(table[j].addr, table[j].port, timeout);
            const std::int32_t me = R;
            write_all(s.sock[j], &me, sizeof(me));
        }
        for (int j = R + 1; j < N; ++j) {
            const int fd = ::accept(my_listen, nullptr, nullptr);
            if (fd < 0) fail("accept()");
            tune(fd);
            std::int32_t who = -1;
            read_all(fd, &who, sizeof(who));
            if (who <= R || who >= N || s.sock[who] != -1) fail("unexpected rank in mesh setup");
            s.sock[who] = fd;
        }
        ::close(my_listen);
    }
}

void raw_send(int dst_world, std::uint64_t gid, std::int32_t tag, const void* buf, std::size_t bytes) {
    state_t& s = S();
    frame_header h{gid, tag, s.world_rank, bytes};
    if (dst_world == s.world_rank) {
        message m;
        m.h = h;
        m.data.assign(static_cast<const char*>(buf), static_cast<const char*>(buf) + bytes);
        s.parked[dst_world].push_back(std::move(m));
        return;
    }
    write_all(s.sock[dst_world], &h, sizeof(h));
    if (bytes) write_all(s.sock[dst_world], buf, bytes);
}

void raw_recv(int src_world, std::uint64_t gid, std::int32_t tag, void* buf, std::size_t bytes) {
    state_t& s = S();
    auto& q = s.parked[src_world];
    for (auto it = q.begin(); it != q.end(); ++it) {
        if (it->h.gid == gid && it->h.tag == tag) {
            if (it->h.bytes != bytes) fail("message size mismatch");
            if (bytes) std::memcpy(buf, it->data.data(), bytes);
            q.erase(it);
            return;
        }
    }
    if (src_world == s.world_rank) fail("receive from self without a matching send");
    for (;;) {
        frame_header h;
        read_all(s.sock[src_world], &h, sizeof(h));
        if (h.gid == gid && h.tag == tag) {
            if (h.bytes != bytes) fail("message size mismatch");
            if (bytes) read_all(s.sock[src_world], buf, bytes);
            return;
        }
        message m;
        m.h = h;
        m.data.resize(h.bytes);
        if (h.bytes) read_all(s.sock[src_world], m.data.data(), h.bytes);
        q.push_back(std::move(m));
    }
}

std::uint64_t mix(std::uint64_t a, std::uint64_t b) {
    std::uint64_t h = a ^ (b + 0x9e3779b97f4a7c15ull + (a << 6) + (a >> 2));
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33;
    return h ? h : 1;
}

template <typename T>
void combine(T* acc, const T* in, int count, op o) {
    for (int i = 0; i < count; ++i) {
        if (o == op::sum) acc[i] = acc[i] + in[i];
        else if (o == op::min) acc[i] = std::min(acc[i], in[i]);
        else acc[i] = std::max(acc[i], in[i]);
    }
}
template <typename T>
void combine_sum_only(T* acc, const T* in, int count, op o) {
    if (o != op::sum) fail("min/max of complex values");
    for (int i = 0; i < count; ++i) acc[i] += in[i];
}

void combine_any(void* acc, const void* in, int count, dtype t, op o) {
    switch (t) {
        case dtype::byte_: case dtype::char_: combine(static_cast<signed char*>(acc), static_cast<const signed char*>(in), count, o); break;
        case dtype::bool_: combine(static_cast<unsigned char*>(acc), static_cast<const unsigned char*>(in), count, o); break;
        case dtype::int_: combine(static_cast<int*>(acc), static_cast<const int*>(in), count, o); break;
        case dtype::long_long_: combine(static_cast<long long*>(acc), static_cast<const long long*>(in), count, o); break;
        case dtype::unsigned_long_long_: combine(static_cast<unsigned long long*>(acc), static_cast<const unsigned long long*>(in), count, o); break;
        case dtype::float_: combine(static_cast<float*>(acc), static_cast<const float*>(in), count, o); break;
        case dtype::double_: combine(static_cast<double*>(acc), static_cast<const double*>(in), count, o); break;
        case dtype::complex_float_: combine_sum_only(static_cast<std::complex<float>*>(acc), static_cast<const std::complex<float>*>(in), count, o); break;
        case dtype::complex_double_: combine_sum_only(static_cast<std::complex<double>*>(acc), static_cast<const std::complex<double>*>(in), count, o); break;
    }
}

void check(const group* g) {
    if (!g) fail("null communicator");
}

}  // namespace

std::size_t dtype_size(dtype t) {
    switch (t) {
        case dtype::byte_: case dtype::char_: case dtype::bool_: return 1;
        case dtype::int_: case dtype::float_: return 4;
        case dtype::long_long_: case dtype::unsigned_long_long_: case dtype::double_: case dtype::complex_float_: return 8;
        case dtype::complex_double_: return 16;
    }
    return 1;
}

void init() {
    state_t& s = S();
    std::lock_guard<std::recursive_mutex> lock(s.mu);
    if (s.up) return;
    s.world_size = std::max(1, env_int("WORLD_SIZE", 1));
    s.world_rank = env_int("RANK", 0);
    s.local_rank = env_int("LOCAL_RANK", s.world_rank);
    if (s.world_rank < 0 || s.world_rank >= s.world_size) fail("RANK outside [0, WORLD_SIZE)");
    connect_world();
    s.world = new group;
    s.world->gid = 1;
    for (int i = 0; i < s.world_size; ++i) s.world->members.push_back(i);
    s.world->my_pos = s.world_rank;
    s.up = true;
}

bool initialized() { return S().up; }

void finalize() {
    state_t& s = S();
    std::lock_guard<std::recursive_mutex> lock(s.mu);
    if (!s.up) return;
    barrier(s.world);
    for (int& fd : s.sock)
        if (fd >= 0) { ::close(fd); fd = -1; }
    delete s.world;
    s.world = nullptr;
    s.up = false;
}

group* world() {
    init();
    return S().world;
}
int local_rank() {
    init();
    return S().local_rank;
}

int rank(const group* g) { check(g); return g->my_pos; }
int size(const group* g) { check(g); return static_cast<int>(g->members.size()); }
std::uint64_t id(const group* g) { check(g); return g->gid; }

void send(group* g, const void* buf, std::size_t bytes, int dst, int tag) {
    check(g);
    std::lock_guard<std::recursive_mutex> lock(S().mu);
    if (dst < 0 || dst >= size(g)) fail("send: destination outside the communicator");
    raw_send(g->members[dst], g->gid, tag, buf, bytes);
}
void recv(group* g, void* buf, std::size_t bytes, int src, int tag) {
    check(g);
    std::lock_guard<std::recursive_mutex> lock(S().mu);
    if (src < 0 || src >= size(g)) fail("recv: source outside the communicator");
    raw_recv(g->members[src], g->gid, tag, buf, bytes);
}

int recv_any(group* g, void* buf, std::size_t bytes, int tag) {
    check(g);
    state_t& s = S();
    std::lock_guard<std::recursive_mutex> lock(s.mu);
    const int n = size(g);
    for (;;) {
        // anything already parked?
        for (int i = 0; i < n; ++i) {
            auto& q = s.parked[g->members[i]];
            for (auto it = q.begin(); it != q.end(); ++it) {
                if (it->h.gid == g->gid && it->h.tag == tag) {
                    if (it->h.bytes != bytes) fail("message size mismatch");
                    if (bytes) std::memcpy(buf, it->data.data(), bytes);
                    q.erase(it);
                    return i;
                }
            }
        }
        // wait until some member's socket has data, and park one frame from every readable socket
        std::vector<pollfd> fds;
        std::vector<int> who;
        for (int i = 0; i < n; ++i) {
            const int w = g->members[i];
            if (w == s.world_rank || s.sock[w] < 0) continue;
            fds.push_back(pollfd{s.sock[w], POLLIN, 0});
            who.push_back(w);
        }
        if (fds.empty()) fail("recv_any: nobody to receive from");
        const int limit = env_int("COSMA_B200_PG_RECV_TIMEOUT", 0);
        const int ready = ::poll(fds.data(), fds.size(), limit > 0 ? limit * 1000 : -1);
        if (ready < 0) {
            if (errno == EINTR) continue;
            fail(std::string("poll: ") + std::strerror(errno));
        }
        if (ready == 0) fail("no message within COSMA_B200_PG_RECV_TIMEOUT seconds: giving up");
        for (size_t f = 0; f < fds.size(); ++f) {
            if (!(fds[f].revents & (POLLIN | POLLHUP))) continue;
            message m;
            read_all(fds[f].fd, &m.h, sizeof(m.h));
            m.data.resize(m.h.bytes);
            if (m.h.bytes) read_all(fds[f].fd, m.data.data(), m.h.bytes);
            s.parked[who[f]].push_back(std::move(m));
        }
    }
}

void bcast(group* g, void* buf, std::size_t bytes, int root) {
    check(g);
    std::lock_guard<std::recursive_mutex> lock(S().mu);
    const int n = size(g);
    if (root < 0 || root >= n) fail("bcast: root outside the communicator");
    if (g->my_pos == root) {
        for (int i = 0; i < n; ++i)
            if (i != root) raw_send(g->members[i], g->gid, TAG_COLL, buf, bytes);
    } else {
        raw_recv(g->members[root], g->gid, TAG_COLL, buf, bytes);
    }
}

void gather(group* g, const void* sendbuf, std::size_t bytes, void* recvbuf, int root) {
    check(g);
    std::lock_guard<std::recursive_mutex> lock(S().mu);
    const int n = size(g);
    if (root < 0 || root >= n) fail("gather: root outside the communicator");
    if (g->my_pos == root) {
        char* out = static_cast<char*>(recvbuf);
        for (int i = 0; i < n; ++i) {
            if (i == root) {
                if (bytes) std::memcpy(out + static_cast<std::size_t>(i) * bytes, sendbuf, bytes);
            } else {
                raw_recv(g->members[i], g->gid, TAG_COLL, out + static_cast<std::size_t>(i) * bytes, bytes);
            }
        }
    } else {
        raw_send(g->members[root], g->gid, TAG_COLL, sendbuf, bytes);
    }
}

void allgather(group* g, const void* sendbuf, std::size_t bytes, void* recvbuf) {
    gather(g, sendbuf, bytes, recvbuf, 0);
    bcast(g, recvbuf, bytes * static_cast<std::size_t>(size(g)), 0);
}

void barrier(group* g) {
    char token = 0;
    std::vector<char> all(static_cast<std::size_t>(size(g)));
    gather(g, &token, 1, all.data(), 0);
    bcast(g, &token, 1, 0);
}

void reduce(group* g, const void* sendbuf, void* recvbuf, int count, dtype t, op o, int root) {
    check(g);
    const std::size_t bytes = dtype_size(t) * static_cast<std::size_t>(count);
    const int n = size(g);
    std::vector<char> all;
    if (g->my_pos == root) all.resize(bytes * static_cast<std::size_t>(n));
    gather(g, sendbuf, bytes, all.data(), root);
    if (g->my_pos == root) {
        // rank order, like a linear MPI reduction: deterministic
        std::vector<char> acc(all.begin(), all.begin() + static_cast<std::ptrdiff_t>(bytes));
        for (int i = 1; i < n; ++i) combine_any(acc.data(), all.data() + static_cast<std::size_t>(i) * bytes, count, t, o);
        if (bytes) std::memcpy(recvbuf, acc.data(), bytes);
    }
}

void allreduce(group* g, const void* sendbuf, void* recvbuf, int count, dtype t, op o) {
    const std::size_t bytes = dtype_size(t) * static_cast<std::size_t>(count);
    std::vector<char> tmp(bytes);
    reduce(g, sendbuf, tmp.data(), count, t, o, 0);
    bcast(g, tmp.data(), bytes, 0);
    if (bytes) std::memcpy(recvbuf, tmp.data(), bytes);
}

group* split(group* g, int color, int key) {
    check(g);
    std::lock_guard<std::recursive_mutex> lock(S().mu);
    const int n = size(g);
    const std::int32_t mine[2] = {color, key};
    std::vector<std::int32_t> all(static_cast<std::size_t>(n) * 2);
    allgather(g, mine, sizeof(mine), all.data());
    const std::uint64_t child = ++g->children;
    if (color < 0) return nullptr;
    std::vector<std::pair<std::pair<int, int>, int>> picked;  // ((key, old rank), world rank)
    for (int i = 0; i < n; ++i)
        if (all[2 * i] == color) picked.push_back({{all[2 * i + 1], i}, g->members[i]});
    std::sort(picked.begin(), picked.end());
    auto* out = new group;
    out->gid = mix(mix(g->gid, child), static_cast<std::uint64_t>(color) + 0x51ull);
    for (std::size_t i = 0; i < picked.size(); ++i) {
        out->members.push_back(picked[i].second);
        if (picked[i].second == S().world_rank) out->my_pos = static_cast<int>(i);
    }
    return out;
}

group* create_group(group* g, const std::vector<int>& ranks, int tag) {
    check(g);
    std::uint64_t h = mix(g->gid, 0x6a09e667f3bcc908ull + static_cast<std::uint64_t>(static_cast<std::uint32_t>(tag)));
    int my_pos = -1;
    for (std::size_t i = 0; i < ranks.size(); ++i) {
        if (ranks[i] < 0 || ranks[i] >= size(g)) fail("create_group: rank outside the parent communicator");
        h = mix(h, static_cast<std::uint64_t>(ranks[i]) + 1);
        if (ranks[i] == g->my_pos) my_pos = static_cast<int>(i);
    }
    if (my_pos < 0) return nullptr;
    auto* out = new group;
    out->gid = h;
    out->my_pos = my_pos;
    for (int r : ranks) out->members.push_back(g->members[r]);
    return out;
}

group* dup(group* g) { return split(g, 0, rank(g)); }

void free(group* g) {
    if (g && g != S().world) delete g;
}

}  // namespace pg
}  // namespace cosma
