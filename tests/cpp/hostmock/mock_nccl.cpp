// TEST INFRASTRUCTURE ONLY — an in-process stand-in for the few NCCL entry points api.cu resolves with dlsym
// (built as build/mock/libnccl.so.2 and found through LD_LIBRARY_PATH by the host-mock driver only).  Every rank of a
// "job" is a THREAD of one process, each with its own zk_ctx; "device" memory is host memory (mock_cudart.cpp), so a
// collective is a rendezvous of the ranks' threads at a barrier followed by plain copies / 64-bit sums.
// Semantics kept from NCCL: ncclAllReduce(sum, uint64) and ncclAllGather may be in place; ncclSend/ncclRecv only take
// effect at ncclGroupEnd, sends and receives between a pair of ranks match in posting order, self send/recv is allowed.
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

struct Posted { int peer; void* ptr; size_t bytes; bool done; };

struct Group {
    int nranks = 0;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    uint64_t generation = 0;
    std::vector<const void*> src;               // per rank: source buffer of the collective in flight
    std::vector<std::vector<Posted>> sends;     // per rank: posted by ncclSend inside the open group
    std::vector<std::vector<uint64_t>> scratch; // per rank: result staged before the in-place write

    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const uint64_t gen = generation;
        if (++waiting == nranks) {
            waiting = 0;
            generation++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen; });
        }
    }
};

struct Comm {
    Group* group;
    int rank;
    std::vector<Posted> recvs;  // posted by ncclRecv inside the open group
    bool in_group = false;
};

std::mutex g_mu;
// never destroyed (the groups stay reachable until the process ends, so leak checkers do not count them)
std::map<std::string, Group*>& g_groups = *new std::map<std::string, Group*>();
uint64_t g_next_id = 1;
thread_local Comm* t_open_comm = nullptr;  // the communicator the calling rank opened a group on

size_t dtype_bytes(int dtype) {
    switch (dtype) {
        case 0: case 1: return 1;   // ncclInt8 / ncclUint8
        case 2: case 3: return 4;   // ncclInt32 / ncclUint32
        case 4: case 5: return 8;   // ncclInt64 / ncclUint64
        default: return 0;
    }
}

}  // namespace

extern "C" {
#define MOCK_API __attribute__((visibility("default")))

struct ncclUniqueId { char internal[128]; };

MOCK_API int ncclGetUniqueId(ncclUniqueId* id) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::memset(id->internal, 0, 128);
    const uint64_t v = g_next_id++;
    std::memcpy(id->internal, "hostmock", 8);
    std::memcpy(id->internal + 8, &v, 8);
    return 0;
}

MOCK_API int ncclCommInitRank(void** comm, int nranks, ncclUniqueId id, int rank) {
    Group* g = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        const std::string key(id.internal, 128);
        auto it = g_groups.find(key);
        if (it == g_groups.end()) {
            g = new Group();
            g->nranks = nranks;
            g->src.assign((size_t)nranks, nullptr);
            g->sends.assign((size_t)nranks, {});
            g->scratch.assign((size_t)nranks, {});
            g_groups[key] = g;
        } else {
            g = it->second;
        }
    }
    if (g->nranks != nranks || rank < 0 || rank >= nranks) return 4;  // ncclInvalidArgument
    *comm = new Comm{g, rank, {}, false};
    g->barrier();  // like NCCL: returns once every rank has joined
    return 0;
}

MOCK_API int ncclCommDestroy(void* comm) {
    delete (Comm*)comm;  // the Group objects live for the life of the process
    return 0;
}

MOCK_API const char* ncclGetErrorString(int) { return "host mock nccl error"; }

MOCK_API int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void*) {
    Comm* c = (Comm*)comm;
    Group* g = c->group;
    if (dtype != 5 || op != 0) return 4;  // only ncclSum over ncclUint64 is used by the library
    g->src[(size_t)c->rank] = send;
    g->barrier();
    std::vector<uint64_t>& acc = g->scratch[(size_t)c->rank];
    acc.assign(count, 0);
    for (int r = 0; r < g->nranks; r++)
        for (size_t i = 0; i < count; i++) acc[i] += ((const uint64_t*)g->src[(size_t)r])[i];
    g->barrier();  // everybody has read every source: in-place results may now be written
    std::memcpy(recv, acc.data(), count * 8);
    g->barrier();
    return 0;
}

MOCK_API int ncclAllGather(const void* send, void* recv, size_t sendcount, int dtype, void* comm, void*) {
    Comm* c = (Comm*)comm;
    Group* g = c->group;
    const size_t bytes = sendcount * dtype_bytes(dtype);
    if (!bytes && sendcount) return 4;
    g->src[(size_t)c->rank] = send;
    g->barrier();
    std::vector<uint64_t>& tmp = g->scratch[(size_t)c->rank];
    tmp.assign(((size_t)g->nranks * bytes + 7) / 8, 0);
    for (int r = 0; r < g->nranks; r++) std::memcpy((char*)tmp.data() + (size_t)r * bytes, g->src[(size_t)r], bytes);
    g->barrier();
    std::memcpy(recv, tmp.data(), (size_t)g->nranks * bytes);
    g->barrier();
    return 0;
}

MOCK_API int ncclGroupStart(void) { return 0; }

MOCK_API int ncclSend(const void* send, size_t count, int dtype, int peer, void* comm, void*) {
    Comm* c = (Comm*)comm;
    if (peer < 0 || peer >= c->group->nranks) return 4;
    t_open_comm = c;
    c->group->sends[(size_t)c->rank].push_back(Posted{peer, const_cast<void*>(send), count * dtype_bytes(dtype), false});
    return 0;
}

MOCK_API int ncclRecv(void* recv, size_t count, int dtype, int peer, void* comm, void*) {
    Comm* c = (Comm*)comm;
    if (peer < 0 || peer >= c->group->nranks) return 4;
    t_open_comm = c;
    c->recvs.push_back(Posted{peer, recv, count * dtype_bytes(dtype), false});
    return 0;
}

MOCK_API int ncclGroupEnd(void) {
    Comm* c = t_open_comm;
    if (!c) return 0;  // an empty group
    t_open_comm = nullptr;
    Group* g = c->group;
    int rc = 0;
    g->barrier();  // every rank has posted
    for (Posted& rv : c->recvs) {
        bool matched = false;
        for (Posted& sd : g->sends[(size_t)rv.peer]) {  // the peer's sends to me, in posting order
            if (sd.peer != c->rank || sd.done) continue;
            if (sd.bytes != rv.bytes) { rc = 4; }
            else std::memcpy(rv.ptr, sd.ptr, rv.bytes);
            sd.done = true;  // only this rank touches the `done` flags of sends addressed to it
            matched = true;
            break;
        }
        if (!matched) rc = 4;
    }
    g->barrier();  // every rank has received
    for (const Posted& sd : g->sends[(size_t)c->rank])
        if (!sd.done) rc = 4;  // a send nobody received
    g->sends[(size_t)c->rank].clear();
    c->recvs.clear();
    g->barrier();
    return rc;
}
}
