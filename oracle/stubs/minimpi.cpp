// TEST INFRASTRUCTURE ONLY -- "minimpi": a small multi-process MPI stand-in used to compile and RUN the UNMODIFIED
// reference sources under /root/reference (there is no MPI in this image; SURVEY.md 8c "Shim 2").
//
// Ranks are processes (the reference keeps per-process singletons: src/cosma/context.cpp:155-158, COSTA
// workspace.hpp:59-63) started by oracle/minirun.py, which creates one unix socketpair per pair of ranks and passes
// the descriptors through MINIMPI_RANK / MINIMPI_SIZE / MINIMPI_FDS. Without those variables the process is rank 0
// of 1 (so libcosma_ref.so also loads in-process through ctypes). Everything is built on one eager, ordered,
// non-blocking point-to-point layer with (context, source, tag) matching; collectives are p2p patterns that reduce
// in rank order (deterministic). RMA entry points exist so one_sided_communicator.cpp links, and abort if reached
// (the reference only uses them with COSMA_OVERLAP_COMM_AND_COMP=ON, default OFF).
// Nothing under cosma_b200/ may include or link this.
#include "mpi.h"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <vector>

#include <fcntl.h>
#include <poll.h>
#include <sys/socket.h>
#include <time.h>
#include <unistd.h>

namespace {

[[noreturn]] void die(const char* what) {
    std::fprintf(stderr, "minimpi[%s]: %s\n", std::getenv("MINIMPI_RANK") ? std::getenv("MINIMPI_RANK") : "0", what);
    std::abort();
}

struct Header {
    int32_t ctx;
    int32_t tag;
    int64_t bytes;
};

struct Message {  // a fully received message nobody asked for yet
    int src;      // world rank
    Header h;
    std::vector<char> data;
};

struct Request {
    bool active = false, is_send = false, done = false;
    // recv side
    void* buf = nullptr;
    int64_t cap = 0;
    int ctx = 0, src = 0 /* world rank or MPI_ANY_SOURCE */, tag = 0;
    const std::vector<int>* comm_world = nullptr;  // to translate the matched source back to a comm rank
    MPI_Status st{};
};

struct Outgoing {
    Header h;
    const char* data;
    int64_t off = 0;  // bytes of header+payload already written
    int req;          // request index to complete
};

struct Incoming {
    Header h;
    int64_t hdr_got = 0, got = 0;
    std::vector<char> data;
};

struct Comm {
    int ctx;
    std::vector<int> world;  // comm rank -> world rank
    int me;                  // my comm rank
};

int g_rank = 0, g_size = 1;
std::vector<int> g_fd;
std::vector<std::deque<Outgoing>> g_out;
std::vector<Incoming> g_in;
std::vector<char> g_closed;  // peer closed its socket
std::deque<Message> g_unexpected;
std::vector<Request> g_req(1);  // index 0 = MPI_REQUEST_NULL
std::vector<int> g_posted;      // posted receives, in order
std::map<int, Comm> g_comm;
std::map<int, std::vector<int>> g_group;
int g_next_handle = 16, g_next_ctx = 16;
bool g_initialized = false, g_finalized = false, g_setup = false;

constexpr int TAG_COLL = -1000;  // internal tags live below every user tag and below MPI_ANY_TAG

size_t tsize(MPI_Datatype t) { return size_t(t & 0xff); }

void setup() {
    if (g_setup) return;
    g_setup = true;
    const char* r = std::getenv("MINIMPI_RANK");
    const char* s = std::getenv("MINIMPI_SIZE");
    const char* f = std::getenv("MINIMPI_FDS");
    if (r && s && f) {
        g_rank = std::atoi(r);
        g_size = std::atoi(s);
        const char* p = f;
        while (*p) {
            g_fd.push_back(std::atoi(p));
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
        if ((int)g_fd.size() != g_size) die("MINIMPI_FDS does not list one descriptor per rank");
        for (int i = 0; i < g_size; ++i)
            if (i != g_rank) {
                int fl = fcntl(g_fd[i], F_GETFL, 0);
                if (fl < 0 || fcntl(g_fd[i], F_SETFL, fl | O_NONBLOCK) < 0) die("fcntl failed on a rank socket");
                int sz = 8 << 20;
                setsockopt(g_fd[i], SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
                setsockopt(g_fd[i], SOL_SOCKET, SO_RCVBUF, &sz, sizeof sz);
            }
    } else {
        g_fd.assign(1, -1);
    }
    g_out.resize(g_size);
    g_in.resize(g_size);
    g_closed.assign(g_size, 0);
    Comm w;
    w.ctx = 1;
    w.me = g_rank;
    for (int i = 0; i < g_size; ++i) w.world.push_back(i);
    g_comm[MPI_COMM_WORLD] = w;
    Comm self;
    self.ctx = 2;
    self.me = 0;
    self.world = {g_rank};
    g_comm[MPI_COMM_SELF] = self;
    g_group[MPI_GROUP_EMPTY] = {};
}

Comm& comm_of(MPI_Comm c) {
    setup();
    auto it = g_comm.find(c);
    if (it == g_comm.end()) die("invalid communicator handle");
    return it->second;
}

int new_request() {
    for (size_t i = 1; i < g_req.size(); ++i)
        if (!g_req[i].active) {
            g_req[i] = Request();
            g_req[i].active = true;
            return (int)i;
        }
    g_req.emplace_back();
    g_req.back().active = true;
    return (int)g_req.size() - 1;
}

bool matches(const Request& q, int src, const Header& h) {
    if (q.ctx != h.ctx) return false;
    if (q.src != MPI_ANY_SOURCE && q.src != src) return false;
    if (q.tag == MPI_ANY_TAG) return h.tag >= 0;
    return q.tag == h.tag;
}

void complete_recv(Request& q, int src, const Header& h, const char* data) {
    if (h.bytes > q.cap) die("message longer than the receive buffer");
    if (h.bytes) std::memcpy(q.buf, data, (size_t)h.bytes);
    int crank = src;
    if (q.comm_world) {
        auto it = std::find(q.comm_world->begin(), q.comm_world->end(), src);
        crank = int(it - q.comm_world->begin());
    }
    q.st.MPI_SOURCE = crank;
    q.st.MPI_TAG = h.tag;
    q.st.MPI_ERROR = 0;
    q.st.count_ = (int)h.bytes;
    q.done = true;
}

void deliver(int src, const Header& h, std::vector<char>&& data) {
    for (size_t i = 0; i < g_posted.size(); ++i) {
        Request& q = g_req[g_posted[i]];
        if (matches(q, src, h)) {
            complete_recv(q, src, h, data.data());
            g_posted.erase(g_posted.begin() + i);
            return;
        }
    }
    Message m;
    m.src = src;
    m.h = h;
    m.data = std::move(data);
    g_unexpected.push_back(std::move(m));
}

// one round of socket progress; blocks in poll() when `block`
void progress(bool block) {
    if (g_size == 1) return;
    std::vector<pollfd> p;
    p.reserve(g_size);
    std::vector<int> who;
    for (int i = 0; i < g_size; ++i) {
        if (i == g_rank || g_closed[i]) continue;
        pollfd e{};
        e.fd = g_fd[i];
        e.events = POLLIN | (g_out[i].empty() ? 0 : POLLOUT);
        p.push_back(e);
        who.push_back(i);
    }
    if (p.empty()) return;
    int rc = poll(p.data(), p.size(), block ? 1000 : 0);
    if (rc < 0) {
        if (errno == EINTR) return;
        die("poll failed");
    }
    for (size_t j = 0; j < p.size(); ++j) {
        const int peer = who[j];
        if (p[j].revents & POLLOUT) {
            auto& q = g_out[peer];
            while (!q.empty()) {
                Outgoing& o = q.front();
                const int64_t total = (int64_t)sizeof(Header) + o.h.bytes;
                ssize_t w;
                if (o.off < (int64_t)sizeof(Header))
                    w = send(p[j].fd, reinterpret_cast<const char*>(&o.h) + o.off, sizeof(Header) - o.off, MSG_NOSIGNAL);
                else
                    w = send(p[j].fd, o.data + (o.off - sizeof(Header)), size_t(total - o.off), MSG_NOSIGNAL);
                if (w < 0) {
                    if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) break;
                    die("send to a peer rank failed (peer exited?)");
                }
                o.off += w;
                if (o.off == total) {
                    g_req[o.req].done = true;
                    q.pop_front();
                }
            }
        }
        if (p[j].revents & (POLLIN | POLLHUP)) {
            Incoming& in = g_in[peer];
            for (;;) {
                ssize_t r;
                if (in.hdr_got < (int64_t)sizeof(Header)) {
                    r = recv(p[j].fd, reinterpret_cast<char*>(&in.h) + in.hdr_got, sizeof(Header) - in.hdr_got, 0);
                    if (r > 0) {
                        in.hdr_got += r;
                        if (in.hdr_got == (int64_t)sizeof(Header)) {
                            in.data.resize((size_t)in.h.bytes);
                            in.got = 0;
                        }
                    }
                } else {
                    r = in.h.bytes > in.got ? recv(p[j].fd, in.data.data() + in.got, size_t(in.h.bytes - in.got), 0) : 1;
                    if (r > 0 && in.h.bytes > in.got) in.got += r;
                }
                if (r == 0) {  // orderly shutdown of the peer: legal once it has passed MPI_Finalize's barrier
                    g_closed[peer] = true;
                    break;
                }
                if (r < 0) {
                    if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) break;
                    die("recv from a peer rank failed");
                }
                if (in.hdr_got == (int64_t)sizeof(Header) && in.got == in.h.bytes) {
                    Header h = in.h;
                    std::vector<char> d;
                    d.swap(in.data);
                    in.hdr_got = 0;
                    in.got = 0;
                    deliver(peer, h, std::move(d));
                }
            }
        }
    }
}

int isend_world(const void* buf, int64_t bytes, int dst_world, int ctx, int tag) {
    const int r = new_request();
    g_req[r].is_send = true;
    Header h{ctx, tag, bytes};
    if (dst_world == g_rank) {
        std::vector<char> d((const char*)buf, (const char*)buf + bytes);
        deliver(g_rank, h, std::move(d));
        g_req[r].done = true;
        return r;
    }
    Outgoing o;
    o.h = h;
    o.data = static_cast<const char*>(buf);
    o.req = r;
    g_out[dst_world].push_back(o);
    progress(false);
    return r;
}

int irecv_world(void* buf, int64_t cap, int src_world, int ctx, int tag, const std::vector<int>* cw) {
    const int r = new_request();
    Request& q = g_req[r];
    q.buf = buf;
    q.cap = cap;
    q.ctx = ctx;
    q.src = src_world;
    q.tag = tag;
    q.comm_world = cw;
    for (auto it = g_unexpected.begin(); it != g_unexpected.end(); ++it)
        if (matches(q, it->src, it->h)) {
            complete_recv(q, it->src, it->h, it->data.data());
            g_unexpected.erase(it);
            return r;
        }
    g_posted.push_back(r);
    return r;
}

void wait_req(int r, MPI_Status* st) {
    if (r == 0) return;
    while (!g_req[r].done) {
        progress(true);
        if (!g_req[r].done && !g_req[r].is_send) {  // never wait on a rank that is gone
            const int src = g_req[r].src;
            bool alive = false;
            for (int i = 0; i < g_size; ++i) alive |= (i != g_rank && !g_closed[i] && (src == MPI_ANY_SOURCE || src == i));
            if (!alive && src != g_rank) die("waiting for a message from a rank that has exited");
        }
    }
    if (st && !g_req[r].is_send) *st = g_req[r].st;
    g_req[r].active = false;
}

// ---- collectives over an explicit rank list (world ranks), so that Comm_create_group can use them too ----
struct Ring {
    int ctx;
    const std::vector<int>& world;
    int me;
    int n() const { return (int)world.size(); }
};

void coll_allgatherv(const Ring& g, const void* send, int64_t sbytes, char* recv, const int64_t* rbytes, const int64_t* roff, int tag) {
    std::vector<int> reqs;
    for (int i = 0; i < g.n(); ++i)
        if (i != g.me) reqs.push_back(irecv_world(recv + roff[i], rbytes[i], g.world[i], g.ctx, tag, nullptr));
    const char* mine = send == MPI_IN_PLACE ? recv + roff[g.me] : static_cast<const char*>(send);
    for (int i = 0; i < g.n(); ++i)
        if (i != g.me) reqs.push_back(isend_world(mine, sbytes, g.world[i], g.ctx, tag));
    if (send != MPI_IN_PLACE && sbytes) std::memmove(recv + roff[g.me], send, (size_t)sbytes);
    for (int r : reqs) wait_req(r, nullptr);
}

void coll_bcast(const Ring& g, void* buf, int64_t bytes, int root, int tag) {
    if (g.n() == 1) return;
    if (g.me == root) {
        std::vector<int> reqs;
        for (int i = 0; i < g.n(); ++i)
            if (i != root) reqs.push_back(isend_world(buf, bytes, g.world[i], g.ctx, tag));
        for (int r : reqs) wait_req(r, nullptr);
    } else {
        wait_req(irecv_world(buf, bytes, g.world[root], g.ctx, tag, nullptr), nullptr);
    }
}

void coll_barrier(const Ring& g, int tag) {
    char z = 0;
    std::vector<char> all(g.n());
    std::vector<int64_t> rb(g.n(), 1), ro(g.n());
    for (int i = 0; i < g.n(); ++i) ro[i] = i;
    coll_allgatherv(g, &z, 1, all.data(), rb.data(), ro.data(), tag);
}

int coll_max_int(const Ring& g, int v, int tag) {
    std::vector<int> all(g.n());
    std::vector<int64_t> rb(g.n(), sizeof(int)), ro(g.n());
    for (int i = 0; i < g.n(); ++i) ro[i] = int64_t(i) * sizeof(int);
    coll_allgatherv(g, &v, sizeof(int), reinterpret_cast<char*>(all.data()), rb.data(), ro.data(), tag);
    return *std::max_element(all.begin(), all.end());
}

template <typename T>
void combine_t(T* acc, const T* in, int64_t n, MPI_Op op) {
    if (op == MPI_SUM) for (int64_t i = 0; i < n; ++i) acc[i] += in[i];
    else if (op == MPI_MIN) for (int64_t i = 0; i < n; ++i) acc[i] = std::min(acc[i], in[i]);
    else if (op == MPI_MAX) for (int64_t i = 0; i < n; ++i) acc[i] = std::max(acc[i], in[i]);
    else die("unsupported reduction op");
}

// acc = acc (op) in for `count` elements of datatype t
void combine(void* acc, const void* in, int64_t count, MPI_Datatype t, MPI_Op op) {
    switch (t) {
        case MPI_CHAR: combine_t((signed char*)acc, (const signed char*)in, count, op); break;
        case MPI_UNSIGNED_CHAR: case MPI_BYTE: case MPI_C_BOOL: combine_t((unsigned char*)acc, (const unsigned char*)in, count, op); break;
        case MPI_SHORT: combine_t((short*)acc, (const short*)in, count, op); break;
        case MPI_INT: combine_t((int*)acc, (const int*)in, count, op); break;
        case MPI_UINT32_T: case MPI_UNSIGNED: combine_t((unsigned*)acc, (const unsigned*)in, count, op); break;
        case MPI_FLOAT: combine_t((float*)acc, (const float*)in, count, op); break;
        case MPI_LONG: case MPI_LONG_LONG: combine_t((long long*)acc, (const long long*)in, count, op); break;
        case MPI_UNSIGNED_LONG: case MPI_UNSIGNED_LONG_LONG: combine_t((unsigned long long*)acc, (const unsigned long long*)in, count, op); break;
        case MPI_DOUBLE: combine_t((double*)acc, (const double*)in, count, op); break;
        case MPI_C_FLOAT_COMPLEX:
            if (op != MPI_SUM) die("complex reductions support MPI_SUM only");
            combine_t((float*)acc, (const float*)in, 2 * count, op);
            break;
        case MPI_C_DOUBLE_COMPLEX:
            if (op != MPI_SUM) die("complex reductions support MPI_SUM only");
            combine_t((double*)acc, (const double*)in, 2 * count, op);
            break;
        default: die("unsupported datatype in a reduction");
    }
}

// out[0..counts[me]) = reduction over ranks (in rank order 0..n-1) of their slice `me` of `in`
void coll_reduce_scatter(const Ring& g, const char* in, char* out, const int64_t* counts, MPI_Datatype t, MPI_Op op, int tag) {
    const size_t ts = tsize(t);
    std::vector<int64_t> off(g.n() + 1, 0);
    for (int i = 0; i < g.n(); ++i) off[i + 1] = off[i] + counts[i];
    const int64_t mine = counts[g.me];
    std::vector<std::vector<char>> parts(g.n());
    std::vector<int> reqs;
    for (int i = 0; i < g.n(); ++i)
        if (i != g.me) {
            parts[i].resize(size_t(mine) * ts);
            reqs.push_back(irecv_world(parts[i].data(), mine * (int64_t)ts, g.world[i], g.ctx, tag, nullptr));
        }
    for (int i = 0; i < g.n(); ++i)
        if (i != g.me) reqs.push_back(isend_world(in + off[i] * ts, counts[i] * (int64_t)ts, g.world[i], g.ctx, tag));
    parts[g.me].assign(in + off[g.me] * ts, in + (off[g.me] + mine) * ts);
    for (int r : reqs) wait_req(r, nullptr);  // all sends out of `in` are complete: `out` may alias it now
    std::vector<char> acc = parts[0];
    for (int i = 1; i < g.n(); ++i) combine(acc.data(), parts[i].data(), mine, t, op);
    if (mine) std::memcpy(out, acc.data(), acc.size());
}

void coll_reduce(const Ring& g, const char* in, char* out, int64_t count, MPI_Datatype t, MPI_Op op, int root, int tag) {
    const int64_t bytes = count * (int64_t)tsize(t);
    if (g.me != root) {
        wait_req(isend_world(in, bytes, g.world[root], g.ctx, tag), nullptr);
        return;
    }
    std::vector<std::vector<char>> parts(g.n());
    std::vector<int> reqs;
    for (int i = 0; i < g.n(); ++i) {
        if (i == root) {
            parts[i].assign(in, in + bytes);
        } else {
            parts[i].resize((size_t)bytes);
            reqs.push_back(irecv_world(parts[i].data(), bytes, g.world[i], g.ctx, tag, nullptr));
        }
    }
    for (int r : reqs) wait_req(r, nullptr);
    std::vector<char> acc = parts[0];
    for (int i = 1; i < g.n(); ++i) combine(acc.data(), parts[i].data(), count, t, op);
    if (bytes) std::memcpy(out, acc.data(), (size_t)bytes);
}

Ring ring_of(const Comm& c) { return Ring{c.ctx, c.world, c.me}; }

MPI_Comm make_comm(const std::vector<int>& world, int ctx) {
    Comm c;
    c.ctx = ctx;
    c.world = world;
    c.me = int(std::find(world.begin(), world.end(), g_rank) - world.begin());
    const int h = g_next_handle++;
    g_comm[h] = c;
    return h;
}

// agree on a fresh context id among `members` (world ranks) using parent context `pctx`
int fresh_ctx(const std::vector<int>& members, int pctx, int tag) {
    const int me = int(std::find(members.begin(), members.end(), g_rank) - members.begin());
    Ring g{pctx, members, me};
    const int id = coll_max_int(g, g_next_ctx, tag);
    g_next_ctx = id + 1;
    return id;
}

void unsupported(const char* what) {
    std::fprintf(stderr, "minimpi: %s is not supported\n", what);
    std::abort();
}

}  // namespace

extern "C" {

int MPI_Init(int*, char***) { setup(); g_initialized = true; return 0; }
int MPI_Init_thread(int*, char***, int req, int* prov) { setup(); g_initialized = true; if (prov) *prov = req; return 0; }
int MPI_Finalize(void) {
    setup();
    if (g_size > 1) {
        Comm& w = comm_of(MPI_COMM_WORLD);
        coll_barrier(ring_of(w), TAG_COLL - 9);
    }
    g_finalized = true;
    return 0;
}
int MPI_Finalized(int* f) { *f = g_finalized; return 0; }
int MPI_Initialized(int* f) { *f = g_initialized; return 0; }
int MPI_Abort(MPI_Comm, int code) { _exit(code ? code : 1); }
double MPI_Wtime(void) { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int MPI_Get_processor_name(char* n, int* l) { std::strcpy(n, "minimpi"); *l = 7; return 0; }

int MPI_Comm_rank(MPI_Comm c, int* r) { *r = comm_of(c).me; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { *s = (int)comm_of(c).world.size(); return 0; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) {
    if (c == MPI_COMM_NULL) { *n = MPI_COMM_NULL; return 0; }
    Comm p = comm_of(c);
    *n = make_comm(p.world, fresh_ctx(p.world, p.ctx, TAG_COLL - 1));
    return 0;
}
// communicator attributes: delete callbacks run when the communicator is freed
static std::vector<std::pair<MPI_Comm_delete_attr_function*, void*>> g_keyvals;
static std::map<std::pair<MPI_Comm, int>, void*> g_attrs;
int MPI_Comm_create_keyval(MPI_Comm_copy_attr_function*, MPI_Comm_delete_attr_function* del, int* keyval, void* extra) {
    g_keyvals.push_back({del, extra});
    *keyval = (int)g_keyvals.size() - 1;
    return 0;
}
int MPI_Comm_set_attr(MPI_Comm c, int keyval, void* value) {
    g_attrs[{c, keyval}] = value;
    return 0;
}
int MPI_Comm_free(MPI_Comm* c) {
    for (auto it = g_attrs.begin(); it != g_attrs.end();) {
        if (it->first.first == *c) {
            auto& kv = g_keyvals[it->first.second];
            if (kv.first) kv.first(*c, it->first.second, it->second, kv.second);
            it = g_attrs.erase(it);
        } else {
            ++it;
        }
    }
    if (*c >= 16) g_comm.erase(*c);
    *c = MPI_COMM_NULL;
    return 0;
}
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n) {
    Comm p = comm_of(c);
    const int np = (int)p.world.size();
    int mine[2] = {color, key};
    std::vector<int> all(2 * np);
    std::vector<int64_t> rb(np, 2 * sizeof(int)), ro(np);
    for (int i = 0; i < np; ++i) ro[i] = int64_t(i) * 2 * sizeof(int);
    coll_allgatherv(ring_of(p), mine, 2 * sizeof(int), reinterpret_cast<char*>(all.data()), rb.data(), ro.data(), TAG_COLL - 2);
    const int ctx = fresh_ctx(p.world, p.ctx, TAG_COLL - 1);  // one id for all colours: their members are disjoint
    if (color == MPI_UNDEFINED) { *n = MPI_COMM_NULL; return 0; }
    std::vector<std::pair<std::pair<int, int>, int>> members;  // ((key, parent rank), world)
    for (int i = 0; i < np; ++i)
        if (all[2 * i] == color) members.push_back({{all[2 * i + 1], i}, p.world[i]});
    std::sort(members.begin(), members.end());
    std::vector<int> world;
    for (auto& m : members) world.push_back(m.second);
    *n = make_comm(world, ctx);
    return 0;
}
int MPI_Comm_split_type(MPI_Comm c, int, int key, MPI_Info, MPI_Comm* n) { return MPI_Comm_split(c, 0, key, n); }
int MPI_Comm_group(MPI_Comm c, MPI_Group* g) {
    const int h = g_next_handle++;
    g_group[h] = comm_of(c).world;
    *g = h;
    return 0;
}
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* n) {
    Comm p = comm_of(c);
    const int ctx = fresh_ctx(p.world, p.ctx, TAG_COLL - 1);
    const std::vector<int>& members = g_group.at(g);
    if (std::find(members.begin(), members.end(), g_rank) == members.end()) { *n = MPI_COMM_NULL; return 0; }
    *n = make_comm(members, ctx);
    return 0;
}
int MPI_Comm_create_group(MPI_Comm c, MPI_Group g, int tag, MPI_Comm* n) {
    Comm p = comm_of(c);
    const std::vector<int> members = g_group.at(g);
    if (std::find(members.begin(), members.end(), g_rank) == members.end()) { *n = MPI_COMM_NULL; return 0; }
    *n = make_comm(members, fresh_ctx(members, p.ctx, TAG_COLL - 100 - tag));
    return 0;
}
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* r) {
    if (a == b) { *r = MPI_IDENT; return 0; }
    const Comm &x = comm_of(a), &y = comm_of(b);
    if (x.world == y.world) { *r = MPI_CONGRUENT; return 0; }
    std::vector<int> u = x.world, v = y.world;
    std::sort(u.begin(), u.end());
    std::sort(v.begin(), v.end());
    *r = u == v ? MPI_SIMILAR : MPI_UNEQUAL;
    return 0;
}
int MPI_Dist_graph_create(MPI_Comm c, int, const int[], const int[], const int[], const int[], MPI_Info, int, MPI_Comm* o) {
    return MPI_Comm_dup(c, o);  // no reordering
}

static MPI_Group new_group(std::vector<int> v) {
    if (v.empty()) return MPI_GROUP_EMPTY;
    const int h = g_next_handle++;
    g_group[h] = std::move(v);
    return h;
}
int MPI_Group_incl(MPI_Group g, int n, const int r[], MPI_Group* o) {
    setup();
    const auto& src = g_group.at(g);
    std::vector<int> v;
    for (int i = 0; i < n; ++i) v.push_back(src.at(r[i]));
    *o = new_group(v);
    return 0;
}
int MPI_Group_excl(MPI_Group g, int n, const int r[], MPI_Group* o) {
    setup();
    const auto& src = g_group.at(g);
    std::vector<int> v;
    for (int i = 0; i < (int)src.size(); ++i)
        if (std::find(r, r + n, i) == r + n) v.push_back(src[i]);
    *o = new_group(v);
    return 0;
}
int MPI_Group_free(MPI_Group* g) {
    if (*g >= 16) g_group.erase(*g);
    *g = MPI_GROUP_NULL;
    return 0;
}
int MPI_Group_union(MPI_Group a, MPI_Group b, MPI_Group* o) {
    setup();
    std::vector<int> v = g_group.at(a);
    for (int x : g_group.at(b))
        if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x);
    *o = new_group(v);
    return 0;
}
int MPI_Group_intersection(MPI_Group a, MPI_Group b, MPI_Group* o) {
    setup();
    const auto& y = g_group.at(b);
    std::vector<int> v;
    for (int x : g_group.at(a))
        if (std::find(y.begin(), y.end(), x) != y.end()) v.push_back(x);
    *o = new_group(v);
    return 0;
}
int MPI_Group_compare(MPI_Group a, MPI_Group b, int* r) {
    setup();
    std::vector<int> u = g_group.at(a), v = g_group.at(b);
    if (u == v) { *r = MPI_IDENT; return 0; }
    std::sort(u.begin(), u.end());
    std::sort(v.begin(), v.end());
    *r = u == v ? MPI_SIMILAR : MPI_UNEQUAL;
    return 0;
}
int MPI_Group_translate_ranks(MPI_Group a, int n, const int r1[], MPI_Group b, int r2[]) {
    setup();
    const auto &x = g_group.at(a), &y = g_group.at(b);
    for (int i = 0; i < n; ++i) {
        auto it = std::find(y.begin(), y.end(), x.at(r1[i]));
        r2[i] = it == y.end() ? MPI_UNDEFINED : int(it - y.begin());
    }
    return 0;
}
int MPI_Group_size(MPI_Group g, int* s) { setup(); *s = (int)g_group.at(g).size(); return 0; }
int MPI_Group_rank(MPI_Group g, int* r) {
    setup();
    const auto& v = g_group.at(g);
    auto it = std::find(v.begin(), v.end(), g_rank);
    *r = it == v.end() ? MPI_UNDEFINED : int(it - v.begin());
    return 0;
}

int MPI_Barrier(MPI_Comm c) { coll_barrier(ring_of(comm_of(c)), TAG_COLL - 3); return 0; }
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) {
    coll_bcast(ring_of(comm_of(c)), b, int64_t(n) * tsize(t), root, TAG_COLL - 4);
    return 0;
}
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
    Comm& p = comm_of(c);
    const int np = (int)p.world.size();
    const int64_t each = int64_t(rn) * tsize(rt);
    std::vector<int64_t> rb(np, each), ro(np);
    for (int i = 0; i < np; ++i) ro[i] = i * each;
    coll_allgatherv(ring_of(p), s, s == MPI_IN_PLACE ? each : int64_t(sn) * tsize(st), static_cast<char*>(r), rb.data(), ro.data(), TAG_COLL - 5);
    return 0;
}
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int rc[], const int d[], MPI_Datatype rt, MPI_Comm c) {
    Comm& p = comm_of(c);
    const int np = (int)p.world.size();
    std::vector<int64_t> rb(np), ro(np);
    for (int i = 0; i < np; ++i) {
        rb[i] = int64_t(rc[i]) * tsize(rt);
        ro[i] = int64_t(d[i]) * tsize(rt);
    }
    coll_allgatherv(ring_of(p), s, s == MPI_IN_PLACE ? rb[p.me] : int64_t(sn) * tsize(st), static_cast<char*>(r), rb.data(), ro.data(), TAG_COLL - 5);
    return 0;
}
int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
    Comm& p = comm_of(c);
    const int np = (int)p.world.size();
    const int64_t sb = int64_t(sn) * tsize(st);
    if (p.me != root) {
        wait_req(isend_world(s, sb, p.world[root], p.ctx, TAG_COLL - 6), nullptr);
        return 0;
    }
    const int64_t each = int64_t(rn) * tsize(rt);
    std::vector<int> reqs;
    for (int i = 0; i < np; ++i) {
        if (i == root) { if (s != MPI_IN_PLACE) std::memmove((char*)r + i * each, s, (size_t)sb); }
        else reqs.push_back(irecv_world((char*)r + i * each, each, p.world[i], p.ctx, TAG_COLL - 6, nullptr));
    }
    for (int q : reqs) wait_req(q, nullptr);
    return 0;
}
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
    Comm& p = comm_of(c);
    const char* in = s == MPI_IN_PLACE ? static_cast<const char*>(r) : static_cast<const char*>(s);
    coll_reduce(ring_of(p), in, static_cast<char*>(r), n, t, op, root, TAG_COLL - 7);
    return 0;
}
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    Comm& p = comm_of(c);
    const char* in = s == MPI_IN_PLACE ? static_cast<const char*>(r) : static_cast<const char*>(s);
    coll_reduce(ring_of(p), in, static_cast<char*>(r), n, t, op, 0, TAG_COLL - 7);
    coll_bcast(ring_of(p), r, int64_t(n) * tsize(t), 0, TAG_COLL - 4);
    return 0;
}
int MPI_Reduce_scatter(const void* s, void* r, const int rc[], MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    Comm& p = comm_of(c);
    std::vector<int64_t> counts(rc, rc + p.world.size());
    const char* in = s == MPI_IN_PLACE ? static_cast<const char*>(r) : static_cast<const char*>(s);
    coll_reduce_scatter(ring_of(p), in, static_cast<char*>(r), counts.data(), t, op, TAG_COLL - 8);
    return 0;
}
int MPI_Reduce_scatter_block(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    Comm& p = comm_of(c);
    std::vector<int64_t> counts(p.world.size(), n);
    const char* in = s == MPI_IN_PLACE ? static_cast<const char*>(r) : static_cast<const char*>(s);
    coll_reduce_scatter(ring_of(p), in, static_cast<char*>(r), counts.data(), t, op, TAG_COLL - 8);
    return 0;
}

int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) {
    Comm& p = comm_of(c);
    *r = isend_world(b, int64_t(n) * tsize(t), p.world.at(d), p.ctx, tag);
    return 0;
}
int MPI_Irecv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request* r) {
    Comm& p = comm_of(c);
    *r = irecv_world(b, int64_t(n) * tsize(t), s == MPI_ANY_SOURCE ? MPI_ANY_SOURCE : p.world.at(s), p.ctx, tag, &p.world);
    return 0;
}
int MPI_Wait(MPI_Request* r, MPI_Status* s) { wait_req(*r, s); *r = MPI_REQUEST_NULL; return 0; }
int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) {
    MPI_Request r;
    MPI_Isend(b, n, t, d, tag, c, &r);
    return MPI_Wait(&r, MPI_STATUS_IGNORE);
}
int MPI_Ssend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { return MPI_Send(b, n, t, d, tag, c); }
int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) {
    MPI_Request r;
    MPI_Irecv(b, n, t, s, tag, c, &r);
    return MPI_Wait(&r, st);
}
int MPI_Waitany(int n, MPI_Request r[], int* idx, MPI_Status* s) {
    bool any = false;
    for (int i = 0; i < n; ++i) any |= r[i] != MPI_REQUEST_NULL;
    if (!any) { *idx = MPI_UNDEFINED; return 0; }
    for (;;) {
        for (int i = 0; i < n; ++i)
            if (r[i] != MPI_REQUEST_NULL && g_req[r[i]].done) {
                wait_req(r[i], s);
                r[i] = MPI_REQUEST_NULL;
                *idx = i;
                return 0;
            }
        progress(true);
    }
}
int MPI_Waitall(int n, MPI_Request r[], MPI_Status s[]) {
    for (int i = 0; i < n; ++i) {
        wait_req(r[i], s ? &s[i] : nullptr);
        r[i] = MPI_REQUEST_NULL;
    }
    return 0;
}
int MPI_Test(MPI_Request* r, int* f, MPI_Status* s) {
    if (*r == MPI_REQUEST_NULL) { *f = 1; return 0; }
    progress(false);
    *f = g_req[*r].done;
    if (*f) { wait_req(*r, s); *r = MPI_REQUEST_NULL; }
    return 0;
}
int MPI_Startall(int, MPI_Request[]) { unsupported("MPI_Startall"); return 1; }
int MPI_Probe(int, int, MPI_Comm, MPI_Status*) { unsupported("MPI_Probe"); return 1; }
int MPI_Get_count(const MPI_Status* s, MPI_Datatype t, int* n) { *n = s ? int(s->count_ / (int)tsize(t)) : 0; return 0; }
int MPI_Get_elements(const MPI_Status* s, MPI_Datatype t, int* n) { return MPI_Get_count(s, t, n); }

int MPI_Info_create(MPI_Info* i) { *i = g_next_handle++; return 0; }
int MPI_Info_set(MPI_Info, const char*, const char*) { return 0; }
int MPI_Info_free(MPI_Info* i) { *i = MPI_INFO_NULL; return 0; }

int MPI_Win_create(void*, MPI_Aint, int, MPI_Info, MPI_Comm, MPI_Win* w) { *w = g_next_handle++; return 0; }
int MPI_Win_free(MPI_Win* w) { *w = MPI_WIN_NULL; return 0; }
int MPI_Win_fence(int, MPI_Win) { return 0; }
int MPI_Win_lock(int, int, int, MPI_Win) { return 0; }
int MPI_Win_unlock(int, MPI_Win) { return 0; }
int MPI_Win_lock_all(int, MPI_Win) { return 0; }
int MPI_Win_unlock_all(MPI_Win) { return 0; }
int MPI_Win_flush_local(int, MPI_Win) { return 0; }
int MPI_Get(void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win) { unsupported("MPI_Get"); return 1; }
int MPI_Rget(void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win, MPI_Request*) { unsupported("MPI_Rget"); return 1; }
int MPI_Accumulate(const void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Op, MPI_Win) { unsupported("MPI_Accumulate"); return 1; }

}  // extern "C"
